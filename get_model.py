"""Repo-root shim so ``from get_model import Model`` works exactly as in the reference tree."""
from image2video_synthesis_using_cinns_b200.get_model import Model  # noqa: F401
