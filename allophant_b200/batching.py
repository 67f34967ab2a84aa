"""Re-exports the batch containers under the reference's module name (``allophant/batching.py``)."""
from .dataset_processing import Batch, LabeledBatch  # noqa: F401
