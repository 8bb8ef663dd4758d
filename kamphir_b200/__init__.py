"""kamphir_b200 -- B200-native phyloK2 tree-kernel hot path.

`PhyloKernel` mirrors the reference class (phyloK2.py:9-236) over libphylok.so (include/phylok.h); the CUDA library
is required -- importing the package is cheap, the first kernel call loads the library and creates the GPU context.
"""
from .phylokernel import PhyloKernel, FlatTree, Forest, Context, get_context, algorithmic_bytes  # noqa: F401
from ._lib import PhyloKError, PK_PREP_KAMPHIR, PK_PREP_LADDERIZE, PK_PREP_ZERO_ROOT  # noqa: F401

__all__ = ["PhyloKernel", "FlatTree", "Forest", "Context", "get_context", "PhyloKError"]
