"""Data path (reference datasets/__init__.py:9-34).  `_name` in the yaml `dataset` block selects the class:
`MelDataset` / `TTSDataset` are the reference's file-backed datasets (own implementation, file_dataset.py),
`SyntheticMelDataset` is the synthetic batch source BASELINE.md section 5 defines.  Unlike the reference the loader
pins its batches (asynchronous H2D) and keeps its workers alive between epochs; `DevicePrefetcher` overlaps the copy of
batch i+1 with step i."""
import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

from .file_dataset import FileDataset, MelDataset, TTSDataset
from .prefetch import DevicePrefetcher


class SyntheticMelDataset(Dataset):
    """mel = clamp(1.5 N(0,1), -4, 4) (B,T,80); wav = clamp(0.3 N(0,1), -1, 1) (300 T, 1); lengths all T"""

    def __init__(self, n_items=1024, n_frames=240, n_mels=80, frameshift=300, seed=1234, feature=None, **_):
        # the reference yaml gives one frameshift per feature (dataset.feature / dataset.frameshift lists,
        # examples/csmsc/configs/msmc_vq_gan.yaml): the mel entry is the hop
        if isinstance(frameshift, (list, tuple)):
            names = list(feature) if feature is not None else []
            frameshift = frameshift[names.index("mel")] if "mel" in names else max(s for s in frameshift if s)
        self.n_items, self.n_frames, self.n_mels, self.seed = n_items, n_frames, n_mels, seed
        self.frameshift = int(frameshift)

    def __len__(self):
        return self.n_items

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        mel = (1.5 * torch.randn(self.n_frames, self.n_mels, generator=g)).clamp_(-4, 4)
        wav = (0.3 * torch.randn(self.n_frames * self.frameshift, 1, generator=g)).clamp_(-1, 1)
        return {"mel": mel, "mel_length": torch.tensor(self.n_frames), "wav": wav,
                "wav_length": torch.tensor(self.n_frames * self.frameshift)}


DATASETS = {"SyntheticMelDataset": SyntheticMelDataset, "MelDataset": MelDataset, "TTSDataset": TTSDataset}


def build_dataset(dataset_config):
    name = dataset_config.get("_name", "SyntheticMelDataset")
    if name not in DATASETS:
        raise ValueError("unknown dataset %s (have: %s)" % (name, ", ".join(sorted(DATASETS))))
    kwargs = {k: v for k, v in dataset_config.items() if not k.startswith("_")}
    return DATASETS[name](**kwargs)


def build_dataloader(dataset_config, dataloader_config, distributed=False):
    ds = build_dataset(dataset_config)
    sampler = DistributedSampler(ds) if distributed else None
    workers = int(dataloader_config.get("num_workers", 0))
    collate = getattr(ds, "collate_fn", None)
    loader = DataLoader(ds, batch_size=dataloader_config.batch_size, shuffle=sampler is None, sampler=sampler,
                        num_workers=workers, collate_fn=collate, pin_memory=torch.cuda.is_available(), drop_last=True,
                        persistent_workers=workers > 0, prefetch_factor=2 if workers > 0 else None)
    return ds, sampler, loader
