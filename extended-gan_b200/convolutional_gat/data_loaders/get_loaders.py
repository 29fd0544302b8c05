"""``get_loaders`` of ``convolutional_gat/data_loaders/get_loaders.py`` (reference :7-39) for the KNMI dataset."""
from .kmni_data_loader import get_loaders as get_loaders_kmni


def get_loaders(train_batch_size: int, test_batch_size: int, preprocessed_folder: str, device, *, dataset: str = "kmni",
                downsample_size=(256, 256), merge_nodes: bool = False, shuffle=True):
    if dataset == "kmni":
        return get_loaders_kmni(train_batch_size, test_batch_size, preprocessed_folder, device, crop=downsample_size[0],
                                merge_nodes=merge_nodes, shuffle=shuffle)
    raise NotImplementedError(f"dataset {dataset!r}: only the KNMI loader is on the conv-GAT path (SURVEY.md section 8f)")
