from typing import Any


def ifnone(a: Any, b: Any) -> Any:
    """``a`` unless it is None, else ``b`` (reference: retinanet/utilities.py:4-9)."""
    return b if a is None else a
