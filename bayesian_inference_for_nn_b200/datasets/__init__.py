from .Dataset import Dataset, ArrayDataset

__all__ = ["Dataset", "ArrayDataset"]
