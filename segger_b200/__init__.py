"""segger_b200 -- B200-native implementation of segger's GNN hot path.

Importing the package does not need a GPU; calling any op does (there is no CPU fallback).
"""
__version__ = "0.1.0"
