"""Drop-in module for kamphir: put this directory first on sys.path (or copy this file next to kamphir.py) and
`from phyloK2 import PhyloKernel` (kamphir.py:10) resolves to the B200 implementation."""
from kamphir_b200.phylokernel import PhyloKernel  # noqa: F401
