"""``datasets.ucf_dataloader_eval.UCF101DataLoader`` as called by evaluate_ucf101.py:59
(``UCF101DataLoader('validation', [224, 224], 1, file_id=..., use_random_start_frame=False)``): whole synthetic
videos ``(video (F,224,224,3), bbox (F,224,224,1), label)``."""
from datasets._synthetic import SyntheticEvalVideos


class UCF101DataLoader(SyntheticEvalVideos):
    NUM_CLASSES = 24
