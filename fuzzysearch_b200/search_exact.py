"""fuzzysearch.search_exact (search_exact.py:22-89), same names.

The package also exports the FUNCTION as ``fuzzysearch_b200.search_exact`` (the reference keeps it inside this
module only); importing this submodule rebinds that package attribute to the module, so the module is made callable:
``from fuzzysearch_b200 import search_exact`` and ``from fuzzysearch_b200.search_exact import search_exact`` both
give something that searches."""
import sys
import types

from .search import ExactSearch, search_exact

__all__ = ["search_exact", "ExactSearch"]


class _CallableModule(types.ModuleType):
    def __call__(self, *args, **kwargs):
        return search_exact(*args, **kwargs)


sys.modules[__name__].__class__ = _CallableModule
