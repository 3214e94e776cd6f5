"""Deterministic synthetic clips in the reference's sample format (U[0,1) RGB clips, one random
axis-aligned box per clip as the localisation mask, random class)."""
import os

import numpy as np
import torch
from torch.utils.data import Dataset


def default_length(kind: str) -> int:
    """Samples per split; ``B200CAPS_SYNTH_LEN`` = "<labeled>,<unlabeled>,<validation>" overrides."""
    env = os.environ.get("B200CAPS_SYNTH_LEN")
    table = dict(labeled=32, unlabeled=64, validation=16)
    if env:
        vals = [int(v) for v in env.split(",")]
        table = dict(zip(("labeled", "unlabeled", "validation"), (vals + vals[-1:] * 3)[:3]))
    return table[kind]


def split_kind(name: str, file_id) -> str:
    if name != "train":
        return "validation"
    return "unlabeled" if "unlabel" in str(file_id).lower() else "labeled"


def make_clip(index: int, kind: str, num_classes: int, depth: int = 8, size: int = 224, seed: int = 47):
    """(video (3,T,H,W) in [0,1), mask (1,T,H,W) in {0,1}, class id) -- a pure function of (index, kind, seed)."""
    salt = dict(labeled=0, unlabeled=1, validation=2)[kind]
    g = torch.Generator().manual_seed(seed * 1000003 + salt * 7919 + index)
    video = torch.rand((3, depth, size, size), generator=g, dtype=torch.float32)
    label = int(torch.randint(0, num_classes, (1,), generator=g))
    y0, x0 = [int(v) for v in torch.randint(0, size - 74, (2,), generator=g)]
    hh, ww = [int(v) for v in torch.randint(30, 74, (2,), generator=g)]
    mask = torch.zeros((1, depth, size, size), dtype=torch.float32)
    mask[:, :, y0:y0 + hh, x0:x0 + ww] = 1.0
    return video, mask, label


class SyntheticTrainClips(Dataset):
    """Training / validation sample dicts: data, aug_data (horizontal flip), loc_msk, action, label_vid."""
    NUM_CLASSES = 24
    WITH_LABEL_VID = True

    def __init__(self, name, clip_shape, file_id, use_random_start_frame=False):
        self.name = "train" if name == "train" else "test"
        self.kind = split_kind(name, file_id)
        self._height, self._width = int(clip_shape[0]), int(clip_shape[1])
        self._size = default_length(self.kind)
        self.vid_files = [(f"synthetic_{self.kind}_{i:05d}", None) for i in range(self._size)]
        self.indexes = np.arange(self._size)
        print(f"[b200caps] {type(self).__module__}: {self._size} SYNTHETIC {self.kind} clips "
              f"(the real videos / pickle splits are not part of this environment)")

    def __len__(self):
        return self._size

    def __getitem__(self, index):
        video, mask, label = make_clip(int(index), self.kind, self.NUM_CLASSES, 8, self._height)
        sample = {"data": video, "loc_msk": mask, "action": torch.Tensor([label]), "aug_data": torch.flip(video, [3])}
        if self.WITH_LABEL_VID:
            sample["label_vid"] = 1 if self.kind != "unlabeled" else 0
        return sample


class SyntheticEvalVideos(Dataset):
    """Whole-video samples for the evaluation scripts: (video (F,H,W,3), bbox (F,H,W,1), label)."""
    NUM_CLASSES = 24
    FRAMES = 40

    def __init__(self, name, clip_shape, *args, file_id=None, use_random_start_frame=False, **kwargs):
        self._height, self._width = int(clip_shape[0]), int(clip_shape[1])
        self._size = default_length("validation")
        self.vid_files = [(f"synthetic_eval_{i:05d}", None) for i in range(self._size)]

    def __len__(self):
        return self._size

    def _video(self, index):
        clips = [make_clip(int(index) * 16 + j, "validation", self.NUM_CLASSES, 8, self._height) for j in range(self.FRAMES // 8)]
        video = torch.cat([c[0] for c in clips], dim=1).permute(1, 2, 3, 0).contiguous().numpy()      # (F,H,W,3)
        bbox = torch.cat([c[1] for c in clips], dim=1).permute(1, 2, 3, 0).contiguous().numpy()       # (F,H,W,1)
        return video, bbox, clips[0][2]

    def __getitem__(self, index):
        return self._video(index)
