// Points-in-polygons spatial join: which transcripts lie strictly inside which (buffered) cell
// outlines -- the tx-neighbors-bd candidate edges the scoring kernel consumes.
//
// Replaces cuspatial.quadtree_point_in_polygon behind
//   points_in_polygons(..., predicate='contains')   /root/reference/src/segger/geometry/query.py:21-100
//   setup_prediction_graph (shape modes)            /root/reference/src/segger/data/utils/neighbors.py:226-238
// (the polygon buffering itself, geopandas .buffer at neighbors.py:229-230, is host geometry in the reference too).
//
// Layout: points are binned into a uniform grid over the polygons' bounding box (cell ~ a polygon's extent) and
// radix-sorted by cell (stable), so a polygon's bounding box covers a few contiguous ranges of the sorted point
// array (one per grid row).  One CTA per polygon keeps its ring in shared memory and streams those ranges; the
// inside test is the even-odd crossing rule evaluated in fp64 with explicitly rounded operations (the CPU oracle
// performs the same operations in the same order, so decisions agree bit for bit, boundary cases included).
// Two passes (count, exclusive scan, fill) and one stable sort by point id give a deterministic point-major,
// polygon-ascending edge list.
#include "sgb_api_internal.cuh"
#include "sgb_sort.cuh"

namespace sgb {
namespace {

constexpr int kPipThreads = 128;
constexpr int kPipMaxSmemVerts = 1024;          // rings longer than this are read from global memory

struct PipGrid {
  double xmin, ymin, inv_cell;
  int nx, ny;
};

__device__ __forceinline__ int pip_cell(double v, double lo, double inv, int n) {
  const double c = floor((v - lo) * inv);
  if (!(c >= 0.0) || c >= static_cast<double>(n)) return -1;
  return static_cast<int>(c);
}

template <typename T>
__global__ void pip_cell_ids_kernel(const T* __restrict__ pts, int64_t n, PipGrid g, uint32_t* __restrict__ cell) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cx = pip_cell(static_cast<double>(pts[2 * i]), g.xmin, g.inv_cell, g.nx);
  const int cy = pip_cell(static_cast<double>(pts[2 * i + 1]), g.ymin, g.inv_cell, g.ny);
  cell[i] = (cx < 0 || cy < 0) ? static_cast<uint32_t>(g.nx) * g.ny : static_cast<uint32_t>(cy) * g.nx + cx;
}

template <typename T>
__global__ void pip_gather_kernel(const T* __restrict__ pts, const uint32_t* __restrict__ perm, int64_t n,
                                  double2* __restrict__ sorted) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = perm[i];
  sorted[i] = make_double2(static_cast<double>(pts[2 * p]), static_cast<double>(pts[2 * p + 1]));
}

// even-odd rule; an edge (a -> b) is crossed by the ray to +x iff exactly one endpoint is strictly above py and
// px < a.x + (py - a.y) * (b.x - a.x) / (b.y - a.y).  Operation order is part of the contract (see oracle).
__device__ __forceinline__ bool pip_inside(const double2* __restrict__ v, int nv, double px, double py) {
  bool in = false;
  double2 a = v[nv - 1];
  for (int e = 0; e < nv; ++e) {
    const double2 b = v[e];
    if ((a.y > py) != (b.y > py)) {
      const double t = __ddiv_rn(__dmul_rn(__dsub_rn(py, a.y), __dsub_rn(b.x, a.x)), __dsub_rn(b.y, a.y));
      if (px < __dadd_rn(a.x, t)) in = !in;
    }
    a = b;
  }
  return in;
}

// FILL == false: counts[poly] = number of points strictly inside;  FILL == true: writes (point, poly) at offsets[poly] + i
template <bool FILL>
__global__ void __launch_bounds__(kPipThreads)
pip_polygon_kernel(const double2* __restrict__ verts, const int64_t* __restrict__ ring_off, int64_t n_poly,
                   const double2* __restrict__ sorted, const uint32_t* __restrict__ perm, const int32_t* __restrict__ cell_start,
                   PipGrid g, int32_t* __restrict__ counts, const int32_t* __restrict__ offsets, uint32_t* __restrict__ out_point,
                   uint32_t* __restrict__ out_poly) {
  __shared__ double2 sv[kPipMaxSmemVerts];
  __shared__ double sbox[4];
  __shared__ int scount;
  const int64_t poly = blockIdx.x;
  if (poly >= n_poly) return;
  const int64_t v0 = ring_off[poly];
  int nv = static_cast<int>(ring_off[poly + 1] - v0);
  // a closed ring that repeats its first vertex at the end: drop the duplicate (zero-length edge)
  if (nv >= 2) {
    const double2 f = verts[v0], l = verts[v0 + nv - 1];
    if (f.x == l.x && f.y == l.y) --nv;
  }
  const bool in_smem = nv <= kPipMaxSmemVerts;
  if (threadIdx.x == 0) { scount = 0; sbox[0] = INFINITY; sbox[1] = INFINITY; sbox[2] = -INFINITY; sbox[3] = -INFINITY; }
  __syncthreads();
  if (in_smem)
    for (int i = threadIdx.x; i < nv; i += kPipThreads) sv[i] = verts[v0 + i];
  // bounding box: one thread (rings are tens of vertices)
  if (threadIdx.x == 0) {
    for (int i = 0; i < nv; ++i) {
      const double2 p = verts[v0 + i];
      sbox[0] = fmin(sbox[0], p.x); sbox[1] = fmin(sbox[1], p.y);
      sbox[2] = fmax(sbox[2], p.x); sbox[3] = fmax(sbox[3], p.y);
    }
  }
  __syncthreads();
  if (nv < 3) {
    if (!FILL && threadIdx.x == 0) counts[poly] = 0;
    return;
  }
  const double2* v = in_smem ? sv : verts + v0;
  int cx0 = pip_cell(sbox[0], g.xmin, g.inv_cell, g.nx), cx1 = pip_cell(sbox[2], g.xmin, g.inv_cell, g.nx);
  int cy0 = pip_cell(sbox[1], g.ymin, g.inv_cell, g.ny), cy1 = pip_cell(sbox[3], g.ymin, g.inv_cell, g.ny);
  // the grid covers the union of all polygon boxes, so the clamps only guard rounding at the upper edge
  cx0 = cx0 < 0 ? 0 : cx0; cy0 = cy0 < 0 ? 0 : cy0;
  cx1 = cx1 < 0 ? g.nx - 1 : cx1; cy1 = cy1 < 0 ? g.ny - 1 : cy1;
  const int base = FILL ? offsets[poly] : 0;
  int mine = 0;
  for (int cy = cy0; cy <= cy1; ++cy) {
    const int beg = cell_start[static_cast<int64_t>(cy) * g.nx + cx0];
    const int end = cell_start[static_cast<int64_t>(cy) * g.nx + cx1 + 1];
    for (int s = beg + threadIdx.x; s < end; s += kPipThreads) {
      const double2 p = sorted[s];
      if (p.x < sbox[0] || p.x > sbox[2] || p.y < sbox[1] || p.y > sbox[3]) continue;
      if (pip_inside(v, nv, p.x, p.y)) {
        if (FILL) {
          const int slot = atomicAdd(&scount, 1);
          out_point[base + slot] = perm[s];
          out_poly[base + slot] = static_cast<uint32_t>(poly);
        } else {
          ++mine;
        }
      }
    }
  }
  if (!FILL) {
    atomicAdd(&scount, mine);
    __syncthreads();
    if (threadIdx.x == 0) counts[poly] = scount;
  }
}

__global__ void pip_emit_kernel(const uint32_t* __restrict__ point, const uint32_t* __restrict__ poly, int64_t E,
                                int32_t* __restrict__ out, int64_t ld) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= E) return;
  out[i] = static_cast<int32_t>(point[i]);
  out[ld + i] = static_cast<int32_t>(poly[i]);
}

struct PipWs {
  uint32_t *cell, *skey, *perm;
  int32_t* cell_start;
  double2* sorted;
  int32_t *counts, *offsets;
  void* sort_ws;
  size_t sort_bytes, total;
};

PipWs pip_carve(void* ws, int64_t n, int64_t ncell, int64_t n_poly) {
  PipWs w{};
  char* p = static_cast<char*>(ws);
  const size_t nb = align_up(static_cast<size_t>(n > 0 ? n : 1) * 4);
  w.cell = reinterpret_cast<uint32_t*>(p); p += nb;
  w.skey = reinterpret_cast<uint32_t*>(p); p += nb;
  w.perm = reinterpret_cast<uint32_t*>(p); p += nb;
  w.cell_start = reinterpret_cast<int32_t*>(p); p += align_up(static_cast<size_t>(ncell + 2) * 4);
  w.sorted = reinterpret_cast<double2*>(p); p += align_up(static_cast<size_t>(n > 0 ? n : 1) * sizeof(double2));
  w.counts = reinterpret_cast<int32_t*>(p); p += align_up(static_cast<size_t>(n_poly + 1) * 4);
  w.offsets = reinterpret_cast<int32_t*>(p); p += align_up(static_cast<size_t>(n_poly + 2) * 4);
  w.sort_ws = p;
  const size_t s1 = sort_pairs_workspace_bytes(n), s2 = scan_workspace_bytes(n_poly + 1);
  w.sort_bytes = s1 > s2 ? s1 : s2;
  w.total = static_cast<size_t>(p - static_cast<char*>(ws)) + w.sort_bytes;
  return w;
}

PipGrid make_grid(double xmin, double ymin, double cell, int nx, int ny) { return PipGrid{xmin, ymin, 1.0 / cell, nx, ny}; }

}  // namespace
}  // namespace sgb

using namespace sgb;

extern "C" size_t sgb_pip_workspace_bytes(int64_t n_points, int64_t n_poly, int nx, int ny) {
  return pip_carve(nullptr, n_points, static_cast<int64_t>(nx) * ny, n_poly).total;
}

// Pass 1: bins the points, counts the contained points of every polygon, scans.  total[0] (device int32) = number of
// (point, polygon) pairs; the caller reads it back to size the output of sgb_pip_fill (same workspace, untouched between).
extern "C" int sgb_pip_count(const void* points, int points_f64, int64_t n_points, const double* verts, const int64_t* ring_off,
                             int64_t n_poly, double xmin, double ymin, double cell, int nx, int ny, int32_t* total, void* ws,
                             size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(n_points >= 0 && n_points < (int64_t(1) << 31) && n_poly >= 0 && n_poly < (int64_t(1) << 31), SGB_ERR_RANGE,
              "pip_count: size out of range");
  SGB_REQUIRE(nx >= 1 && ny >= 1 && static_cast<int64_t>(nx) * ny < (int64_t(1) << 28) && cell > 0.0, SGB_ERR_ARG, "pip_count: bad grid");
  SGB_REQUIRE(total && ws && (n_points == 0 || points) && (n_poly == 0 || (verts && ring_off)), SGB_ERR_ARG, "pip_count: null tensor");
  SGB_REQUIRE(ws_bytes >= sgb_pip_workspace_bytes(n_points, n_poly, nx, ny), SGB_ERR_WORKSPACE, "pip_count: workspace too small");
  if (n_points == 0 || n_poly == 0) { cudaMemsetAsync(total, 0, sizeof(int32_t), stream); return check_launch("pip_count(empty)"); }
  const int64_t ncell = static_cast<int64_t>(nx) * ny;
  PipWs w = pip_carve(ws, n_points, ncell, n_poly);
  const PipGrid g = make_grid(xmin, ymin, cell, nx, ny);
  const unsigned nb = static_cast<unsigned>(ceil_div(n_points, 256));
  if (points_f64) pip_cell_ids_kernel<double><<<nb, 256, 0, stream>>>(static_cast<const double*>(points), n_points, g, w.cell);
  else pip_cell_ids_kernel<float><<<nb, 256, 0, stream>>>(static_cast<const float*>(points), n_points, g, w.cell);
  int rc = sort_pairs(w.cell, nullptr, w.skey, w.perm, n_points, bits_for(ncell + 1), w.sort_ws, w.sort_bytes, stream, true);
  if (rc != SGB_OK) return rc;
  rc = rowptr_from_sorted(w.skey, n_points, w.cell_start, ncell + 1, stream);     // cell_start[ncell] = first out-of-grid point
  if (rc != SGB_OK) return rc;
  if (points_f64) pip_gather_kernel<double><<<nb, 256, 0, stream>>>(static_cast<const double*>(points), w.perm, n_points, w.sorted);
  else pip_gather_kernel<float><<<nb, 256, 0, stream>>>(static_cast<const float*>(points), w.perm, n_points, w.sorted);
  pip_polygon_kernel<false><<<static_cast<unsigned>(n_poly), kPipThreads, 0, stream>>>(
      reinterpret_cast<const double2*>(verts), ring_off, n_poly, w.sorted, w.perm, w.cell_start, g, w.counts, nullptr, nullptr, nullptr);
  rc = exclusive_scan_i32(w.counts, w.offsets, n_poly, w.sort_ws, w.sort_bytes, stream);
  if (rc != SGB_OK) return rc;
  cudaMemcpyAsync(total, w.offsets + n_poly, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
  return check_launch("pip_count");
}

// Pass 2: edge_index [2, E] int32 (row 0 = point index, row 1 = polygon index), point-major, polygon ascending.
// scratch: 4 * E uint32 (pairs + sorted pairs) + sort workspace for E keys: sgb_pip_fill_scratch_bytes(E).
extern "C" size_t sgb_pip_fill_scratch_bytes(int64_t E) {
  return 4 * align_up(static_cast<size_t>(E > 0 ? E : 1) * 4) + sort_pairs_workspace_bytes(E);
}

extern "C" int sgb_pip_fill(const double* verts, const int64_t* ring_off, int64_t n_points, int64_t n_poly, double xmin, double ymin,
                            double cell, int nx, int ny, int64_t E, int32_t* edge_index, int64_t ld, void* ws, size_t ws_bytes,
                            void* scratch, size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(E >= 0 && E < (int64_t(1) << 31), SGB_ERR_RANGE, "pip_fill: E out of range");
  if (E == 0) return SGB_OK;
  SGB_REQUIRE(verts && ring_off && edge_index && ws && scratch && ld >= E, SGB_ERR_ARG, "pip_fill: null tensor");
  SGB_REQUIRE(ws_bytes >= sgb_pip_workspace_bytes(n_points, n_poly, nx, ny), SGB_ERR_WORKSPACE, "pip_fill: workspace too small");
  SGB_REQUIRE(scratch_bytes >= sgb_pip_fill_scratch_bytes(E), SGB_ERR_WORKSPACE, "pip_fill: scratch too small");
  const int64_t ncell = static_cast<int64_t>(nx) * ny;
  PipWs w = pip_carve(ws, n_points, ncell, n_poly);
  const PipGrid g = make_grid(xmin, ymin, cell, nx, ny);
  char* p = static_cast<char*>(scratch);
  const size_t eb = align_up(static_cast<size_t>(E) * 4);
  uint32_t* pt = reinterpret_cast<uint32_t*>(p); p += eb;
  uint32_t* pl = reinterpret_cast<uint32_t*>(p); p += eb;
  uint32_t* spt = reinterpret_cast<uint32_t*>(p); p += eb;
  uint32_t* spl = reinterpret_cast<uint32_t*>(p); p += eb;
  pip_polygon_kernel<true><<<static_cast<unsigned>(n_poly), kPipThreads, 0, stream>>>(
      reinterpret_cast<const double2*>(verts), ring_off, n_poly, w.sorted, w.perm, w.cell_start, g, nullptr, w.offsets, pt, pl);
  // pairs are grouped by polygon (ascending) with arbitrary order inside a group; a stable sort by point id makes the
  // list point-major with polygons ascending per point -- independent of the order the atomics handed out slots
  int rc = sort_pairs(pt, pl, spt, spl, E, bits_for(n_points), p, sort_pairs_workspace_bytes(E), stream);
  if (rc != SGB_OK) return rc;
  pip_emit_kernel<<<static_cast<unsigned>(ceil_div(E, 256)), 256, 0, stream>>>(spt, spl, E, edge_index, ld);
  return check_launch("pip_fill");
}
