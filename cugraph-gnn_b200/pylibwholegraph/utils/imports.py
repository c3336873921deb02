"""Optional-import helper (same contract as the reference's pylibwholegraph/utils/imports.py:
a missing module becomes a placeholder that raises on first attribute access)."""
import importlib


class MissingModule:
    def __init__(self, mod_name):
        self.name = mod_name

    def __getattr__(self, attr):
        raise RuntimeError(f"This feature requires the '{self.name}' package/module")


def import_optional(mod, default_mod_class=MissingModule):
    try:
        return importlib.import_module(mod)
    except ModuleNotFoundError:
        return default_mod_class(mod_name=mod)
