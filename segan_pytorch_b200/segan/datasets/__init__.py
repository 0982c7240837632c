from .se_dataset import (collate_fn, normalize_wave_minmax, pre_emphasize, de_emphasize,
                         SyntheticSEDataset, DevicePrefetcher, SEDataset)

__all__ = ["collate_fn", "normalize_wave_minmax", "pre_emphasize", "de_emphasize", "SyntheticSEDataset",
           "DevicePrefetcher", "SEDataset"]
