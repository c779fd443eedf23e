"""dsa_b200 — B200-native engine for the hot path of DynamicSparseArrays.jl (PMA / PCSR batched updates, finds, SpMV).

The directory is named after the reference package (``dynamicsparsearrays.jl_b200``); because of the dot it is imported
through the ``dsa_b200`` shim at the repository root (``import dsa_b200``).
"""
from . import _lib
from ._lib import (ArgumentError, BoundsError, CudaError, DsaError, ErrorException, build, declared_symbols, device_count, lib,
                   require_gpu, set_tile_mode)
from .api import (Buffer, CharCodec, DynamicMatrixColView, DynamicSparseMatrix, DynamicSparseVector, SparseVector, addrow, closefillmode,
                  deletecolumn, deletepartition, deleterow, dynamicsparse, dynamicsparsevec, KeyCodec, load_checkpoint, nbpartitions, nnz, PackedCSC, save_checkpoint, shrink_size, to_coo)
