"""CheckpointIO for the fd model (reference: fd/checkpoints.py; a missing file raises FileNotFoundError there)."""
from .._checkpoints import make_checkpoint_io

CheckpointIO = make_checkpoint_io(FileNotFoundError)
