"""Bio.Phylo stand-in package (test infrastructure; see Phylo.py)."""
