"""Drop-in for the `src/data` package the reference's drivers import (`from data import get_split_dataset`)
but the reference tree does not contain."""
from pixel_nerf_multiscale_b200.data import DVRDataset, MultiObjectDataset, SRNDataset, get_split_dataset  # noqa: F401
