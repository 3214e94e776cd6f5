"""``datasets.ucf_dataloader.UCF101DataLoader`` with the reference's constructor (ucf_dataloader.py:38) and sample
dict (:189), serving synthetic clips (see datasets/__init__.py)."""
from datasets._synthetic import SyntheticTrainClips


class UCF101DataLoader(SyntheticTrainClips):
    NUM_CLASSES = 24
    WITH_LABEL_VID = True
