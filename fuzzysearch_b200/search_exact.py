"""fuzzysearch.search_exact (search_exact.py:22-89), same names."""
from .search import ExactSearch, search_exact

__all__ = ["search_exact", "ExactSearch"]
