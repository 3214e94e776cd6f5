"""``datasets.jhmdb_dataloader_eval.JHMDB`` for evaluate_jhmdb.py:62: ``(video, bbox, label, video name)``."""
from datasets._synthetic import SyntheticEvalVideos


class JHMDB(SyntheticEvalVideos):
    NUM_CLASSES = 21

    def __getitem__(self, index):
        video, bbox, label = self._video(index)
        return video, bbox, label, self.vid_files[int(index)][0]
