"""`InvalidShapeError` with the reference's message format (chemprop/exceptions.py:6-12)."""
from typing import Iterable


def _pretty_shape(shape: Iterable[int]) -> str:
    return " x ".join(map(str, shape))


class InvalidShapeError(ValueError):
    def __init__(self, var_name: str, received: Iterable[int], expected: Iterable[int]):
        super().__init__(
            f"arg '{var_name}' has incorrect shape! "
            f"got: `{_pretty_shape(received)}`. expected: `{_pretty_shape(expected)}`"
        )
