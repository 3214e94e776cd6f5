"""``datasets.load_jhmdb_pytorch_multi.JHMDB`` -- imported by the reference's main_jhmdb.py:338 but ABSENT from the
reference repository (it is the module datasets/jhmdb_dataloader.py announces itself as, :65).  Same
constructor (``JHMDB(name, clip_shape, file_id, use_random_start_frame)``) and sample dict (jhmdb_dataloader.py:229:
data, loc_msk, action, mask_cls, aug_data -- no ``label_vid``: main_jhmdb.py:68-70 builds the labeled pattern from
ones / zeros), 21 classes, synthetic clips."""
from datasets._synthetic import SyntheticTrainClips


class JHMDB(SyntheticTrainClips):
    NUM_CLASSES = 21
    WITH_LABEL_VID = False

    def __getitem__(self, index):
        sample = super().__getitem__(index)
        sample["mask_cls"] = sample["loc_msk"].clone()
        return sample
