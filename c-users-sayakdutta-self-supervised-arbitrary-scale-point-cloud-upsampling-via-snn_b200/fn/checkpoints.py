"""CheckpointIO for the fn model (reference: fn/checkpoints.py; a missing file raises FileExistsError there)."""
from .._checkpoints import make_checkpoint_io

CheckpointIO = make_checkpoint_io(FileExistsError)
