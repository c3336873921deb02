"""Node-classification datasets / dataloaders of the WholeGraph examples (role of the reference's
pylibwholegraph/torch/data_loader.py)."""
import numpy as np

from pylibwholegraph.utils.imports import import_optional

torch = import_optional("torch")


class NodeClassificationDataset(torch.utils.data.Dataset):
    def __init__(self, raw_dataset):
        self.dataset = raw_dataset

    def __getitem__(self, index):
        return self.dataset[index]

    def __len__(self):
        return len(self.dataset)


def create_node_classification_datasets(data_and_label: dict):
    """{train,valid,test}_{idx,label} arrays -> three datasets of (node id, int64 label) pairs."""
    out = []
    for split in ("train", "valid", "test"):
        idx, label = data_and_label[split + "_idx"], np.asarray(data_and_label[split + "_label"]).astype(np.int64)
        out.append(NodeClassificationDataset(list(zip(idx, label))))
    return tuple(out)


def get_train_dataloader(train_dataset, batch_size: int, *, replica_id: int = 0, num_replicas: int = 1, num_workers: int = 0):
    sampler = torch.utils.data.distributed.DistributedSampler(train_dataset, num_replicas=num_replicas, rank=replica_id,
                                                              shuffle=True, drop_last=True)
    return torch.utils.data.DataLoader(train_dataset, batch_size=batch_size, num_workers=num_workers, pin_memory=True,
                                       persistent_workers=True if num_workers > 0 else None, sampler=sampler)


def get_valid_test_dataloader(valid_test_dataset, batch_size: int, *, num_workers: int = 0):
    sampler = torch.utils.data.distributed.DistributedSampler(valid_test_dataset, num_replicas=1, rank=0, shuffle=False,
                                                              drop_last=False)
    return torch.utils.data.DataLoader(valid_test_dataset, batch_size=batch_size, num_workers=num_workers, pin_memory=True,
                                       sampler=sampler)
