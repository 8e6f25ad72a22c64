// Internal umbrella: public C ABI + shared helpers.
#pragma once
#include "../../include/segger_b200.h"
#include "sgb_common.cuh"
