"""File-backed datasets with the reference's yaml interface (datasets/base_dataset.py, mel_dataset.py,
tts_dataset.py) -- own implementation, host side only.

A dataset is a list of utterance ids plus, per feature, either a path template (`examples/.../mel/{}.npy`,
`.../wav_24k/{}.wav`) or a "book" file holding every utterance's values (`id|v v v ...`, e.g. phone.txt / dur.txt).
Differences from the reference's loader, all on the host:
  * .npy files are opened memory-mapped, so a random `segment_length` window reads only its own bytes (the reference
    re-implements numpy's header parsing for the same effect, utils/utils.py:20-108);
  * .wav files are read with the standard library `wave` module (PCM 16 / 32 bit) instead of libsndfile;
  * the loader pins every collated batch (DataLoader pin_memory=True; the reference sets it to False,
    datasets/__init__.py:29) so the H2D copy can be asynchronous (datasets/prefetch.py)."""
import math
import os
import random
import wave

import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence

MIN_DATASET_SIZE = 3200          # reference base_dataset.py:22: an epoch is at least this many draws


def read_wav(path, start=0, length=-1, shape_only=False):
    """PCM wav -> float32 (frames, 1) in [-1, 1); `start` / `length` in frames"""
    with wave.open(path, "rb") as f:
        n, ch, width = f.getnframes(), f.getnchannels(), f.getsampwidth()
        if shape_only:
            return (n, ch)
        f.setpos(min(max(start, 0), n))
        count = n - start if length <= 0 else min(length, n - start)
        raw = f.readframes(max(count, 0))
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError("unsupported sample width %d in %s" % (width, path))
    return x.reshape(-1, ch)[:, :1]


def read_array(path, dimension=None, start=0, length=-1, shape_only=False):
    ext = os.path.splitext(path)[-1]
    if ext == ".wav":
        return read_wav(path, start, length, shape_only)
    if ext == ".npy":
        a = np.load(path, mmap_mode="r")
    elif ext == ".pt":
        a = torch.load(path, map_location="cpu").squeeze(0).numpy()
        if dimension is not None and a.shape[0] == dimension:
            a = a.T
    elif ext in (".dat", ".mgc", ".ap"):
        a = np.memmap(path, dtype=np.float32, mode="r").reshape(-1, dimension or 1)
    else:
        raise ValueError("unknown feature file type: %s" % path)
    if shape_only:
        return tuple(a.shape)
    end = None if length <= 0 else start + length
    return np.ascontiguousarray(a[start:end])


def parse_values(text, dimension=None):
    """'1 2 3' or '1_0_1 2_1_0' -> float array (n,) or (n, dimension)"""
    x = np.array(text.replace("_", " ").split(), dtype=np.float64)
    if dimension is not None and dimension > 1:
        x = x.reshape(-1, dimension)
    return x


def read_book(path):
    """`id|values[|values...]` per line -> {id: array | [arrays]}"""
    book = {}
    with open(path) as f:
        for line in f:
            parts = line.rstrip("\n").split("|")
            if len(parts) < 2:
                continue
            cols = []
            for col in parts[1:]:
                items = col.split(" ")
                cols.append(np.array([[float(v) for v in it.split("_")] if "_" in it else float(it) for it in items]))
            book[parts[0]] = cols if len(cols) > 1 else cols[0]
    return book


def align_sequences(seqs, frameshift):
    """trim time sequences with different frame shifts to a common duration (reference utils.py align_features)"""
    if len(seqs) < 2:
        return seqs
    dur = min(v.shape[0] * frameshift[k] for k, v in seqs.items())
    return {k: v[: int(dur // frameshift[k])] for k, v in seqs.items()}


class FileDataset(torch.utils.data.Dataset):
    def __init__(self, id_list, feature, samplerate, dimension, frameshift, feature_path=None, feature_stat=None,
                 padding_value=None, segment_length=-1, pre_load=True, seed=1234, training=True):
        super().__init__()
        self.samplerate, self.feature = samplerate, list(feature)
        self.dimension = {f: d for f, d in zip(feature, dimension) if d > 0}
        self.frameshift = {f: s for f, s in zip(feature, frameshift) if s is not None and s > 0}
        self.padding_value = dict(zip(feature, padding_value)) if padding_value is not None else \
            {f: 0 for f in feature}
        self.segment_length, self.pre_load, self.training = segment_length, pre_load, training
        self.feature_stat = {}
        if feature_stat is not None:
            raise NotImplementedError("feature_stat normalisation is unused by the in-tree configs")
        self.rng = random.Random(seed)
        self.items = {}                    # (id, feature) -> ndarray | path
        self.id_list = self._index(id_list, feature_path)
        if self.training:
            self.rng.shuffle(self.id_list)

    # ------------------------------------------------------------------ indexing
    def _index(self, id_list_file, feature_path):
        if isinstance(id_list_file, (list, tuple)):
            ids = []
            for i, one in enumerate(id_list_file):
                ids += self._index(one, [p[i] for p in feature_path])
            return ids
        with open(id_list_file) as f:
            ids = [tuple(line.split()) for line in f if line.strip()]
        for feat, path in zip(self.feature, feature_path):
            if isinstance(path, str) and os.path.isfile(path):       # a book of all utterances
                book = read_book(path)
                for attrs in ids:
                    key = next(a for a in attrs if a in book)
                    self.items[(attrs, feat)] = np.asarray(book[key])
            else:                                                    # one file per utterance
                for attrs in ids:
                    self.items[(attrs, feat)] = path.format(*attrs)
        if self.pre_load and self.training:
            for key, src in list(self.items.items()):
                if isinstance(src, str):
                    self.items[key] = read_array(src, self.dimension.get(key[1]))
        return ids

    def __len__(self):
        return max(MIN_DATASET_SIZE, len(self.id_list)) if self.training else len(self.id_list)

    def __getitem__(self, index):
        return self.parse_case(index % len(self.id_list))

    # ------------------------------------------------------------------ one utterance
    def parse_case(self, index):
        uid = self.id_list[index]
        srcs = {f: self.items[(uid, f)] for f in self.feature if (uid, f) in self.items}
        seg, seg_start = -1, 0.0
        if self.training and self.segment_length > 0:
            seg = self.segment_length
            ref = max(self.frameshift, key=self.frameshift.get)
            n = srcs[ref].shape[0] if not isinstance(srcs[ref], str) else \
                read_array(srcs[ref], self.dimension.get(ref), shape_only=True)[0]
            last = max(0, n - math.ceil(seg / self.frameshift[ref]))
            seg_start = float(self.rng.randint(0, last) * self.frameshift[ref])
        out = {}
        for f, src in srcs.items():
            start, length = 0, -1
            if f in self.frameshift:
                start, length = int(seg_start / self.frameshift[f]), int(seg / self.frameshift[f])
            if isinstance(src, str):
                x = read_array(src, self.dimension.get(f), start, length) if os.path.isfile(src) else \
                    parse_values(src, self.dimension.get(f))[start: (start + length) if length > 0 else None]
                if 0 in x.shape:
                    raise ValueError("cannot parse %s" % src)
            else:
                x = src[start: (start + length) if length > 0 else None]
            out[f] = x
        seqs = align_sequences({k: v for k, v in out.items() if k in self.frameshift}, self.frameshift)
        out.update(seqs)
        if not self.training:
            out["_id"] = index
        return out

    @staticmethod
    def _tensors(batch):
        def conv(v):
            if isinstance(v, np.ndarray):
                v = np.array(v)                       # own, writable copy (memory-mapped sources are read-only)
                return torch.from_numpy(v).float() if v.dtype.kind == "f" else torch.from_numpy(v)
            return v
        return {name: [conv(item[name]) for item in batch] for name in batch[0]}


class MelDataset(FileDataset):
    """(mel, wav) pairs for the autoencoder / vocoder (reference datasets/mel_dataset.py)"""

    def collate_fn(self, batch):
        feats = self._tensors(batch)
        lengths, order = torch.sort(torch.tensor([x.shape[0] for x in feats["mel"]], dtype=torch.int64),
                                    descending=True)
        out = {}
        for k, v in feats.items():
            v = [v[i] for i in order]
            if k in ("dur", "npw"):
                out[k + "_length"] = torch.tensor([x.shape[0] for x in v], dtype=torch.int32)
                v = [x.squeeze(-1) if x.dim() == 2 else x for x in v]
            if torch.is_tensor(v[0]):
                v = pad_sequence(v, batch_first=True, padding_value=self.padding_value[k]) if v[0].dim() >= 1 \
                    else torch.stack(v)
            elif k == "_id":
                v = list(v)
            out[k] = v
        out["mel_length"] = lengths
        if "wav" in out:
            out["wav_length"] = lengths * int(self.frameshift["mel"])
        return out


class TTSDataset(FileDataset):
    """(text, dur, mel) for the acoustic model (reference datasets/tts_dataset.py)"""

    def parse_case(self, index):
        d = super().parse_case(index)
        if d["text"].ndim == 2 and d["text"].shape[1] == 1:
            d["text"] = d["text"][:, 0]
        n_text = len(d["text"])
        if "dur" in d:
            durs = np.array(d["dur"], dtype=np.float64)
            if durs.ndim == 2:
                durs = durs.squeeze(1)
            assert len(durs) == n_text, "%s: %d durations for %d symbols" % (self.id_list[index], len(durs), n_text)
            if "mel" in d:
                if d["mel"].shape[0] / max(durs.sum(), 1e-9) > 100:      # durations in seconds -> frames
                    durs = durs * self.samplerate / self.frameshift["mel"]
                    for i in range(len(durs)):
                        r = round(durs[i])
                        if i < len(durs) - 1:
                            durs[i + 1] += durs[i] - r
                        durs[i] = r
                shift = d["mel"].shape[0] - durs.sum()
                assert -5 <= shift <= 5, "%s: %d frames vs %d" % (self.id_list[index], d["mel"].shape[0], durs.sum())
                durs[-1] += shift
            d["dur"] = durs
        return d

    def collate_fn(self, batch):
        feats = self._tensors(batch)
        lengths, order = torch.sort(torch.tensor([x.shape[0] for x in feats["text"]], dtype=torch.int64),
                                    descending=True)
        feats = {k: [v[i] for i in order] for k, v in feats.items()}
        out = dict(feats)
        out["text_length"] = lengths
        for name in ("text", "tone", "dur"):
            if name in feats:
                out[name] = pad_sequence(feats[name], batch_first=True, padding_value=self.padding_value[name])
        for name in ("mel", "wav", "pitch", "energy"):
            if name in feats:
                if name in ("mel", "wav"):
                    out[name + "_length"] = torch.tensor([x.shape[0] for x in feats[name]], dtype=torch.float32)
                out[name] = pad_sequence(feats[name], batch_first=True, padding_value=self.padding_value[name])
        return out
