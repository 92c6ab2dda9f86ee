"""Data path.  The reference's file-backed datasets (datasets/, utils.py:20-108) are out of the hot-path scope
(SURVEY 2); this backend ships the synthetic batch source BASELINE.md section 5 defines, with pinned host memory so
the H2D copy is asynchronous.  `_name: SyntheticMelDataset` in the yaml selects it."""
import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler


class SyntheticMelDataset(Dataset):
    """mel = clamp(1.5 N(0,1), -4, 4) (B,T,80); wav = clamp(0.3 N(0,1), -1, 1) (300 T, 1); lengths all T"""

    def __init__(self, n_items=1024, n_frames=240, n_mels=80, frameshift=300, seed=1234, feature=None, **_):
        # the reference yaml gives one frameshift per feature (dataset.feature / dataset.frameshift lists,
        # examples/csmsc/configs/msmc_vq_gan.yaml): the mel entry is the hop
        if isinstance(frameshift, (list, tuple)):
            names = list(feature) if feature is not None else []
            frameshift = frameshift[names.index("mel")] if "mel" in names else max(frameshift)
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


def build_dataloader(dataset_config, dataloader_config, distributed=False):
    name = dataset_config.get("_name", "SyntheticMelDataset")
    if name != "SyntheticMelDataset":
        raise NotImplementedError(
            "dataset %s: file-backed datasets are outside this backend's scope; use the reference's "
            "msmctts.datasets for real data, or _name: SyntheticMelDataset" % name)
    kwargs = {k: v for k, v in dataset_config.items() if not k.startswith("_")}
    ds = SyntheticMelDataset(**kwargs)
    sampler = DistributedSampler(ds) if distributed else None
    loader = DataLoader(ds, batch_size=dataloader_config.batch_size, shuffle=sampler is None, sampler=sampler,
                        num_workers=dataloader_config.get("num_workers", 0), pin_memory=True, drop_last=True)
    return ds, sampler, loader
