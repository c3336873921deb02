"""Optional imports (same contract as the reference's cugraph_pyg/utils/imports.py)."""
from pylibwholegraph.utils.imports import MissingModule, import_optional  # noqa: F401
