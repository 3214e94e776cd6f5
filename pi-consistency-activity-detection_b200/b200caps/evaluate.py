"""The evaluation consumer of the hot path: evaluate_ucf101.py:95-187 / evaluate_jhmdb.py (whole videos cut into 8-frame
clips, batches of 14 clips through the eval-mode CapsNet, sigmoid >= 0.5 masks against the ground truth, per-class
frame / video IoU counts over 21 thresholds -> f-mAP / v-mAP, and clip-averaged classification accuracy).

The model runs with BatchNorm folded into the convolutions (engine.EVAL_FOLD_BN) and the collapsed decoder tail; the
per-frame intersection / union counts are taken on the device (b2c_frame_iou_counts), so only 3 integers per frame and
the class activations cross PCIe instead of the (B,1,8,224,224) probability maps the reference copies to the host."""
from __future__ import annotations

from typing import Iterable, Tuple

import numpy as np
import torch

from . import ops


def video_to_clips(video: np.ndarray, bbox: np.ndarray, depth: int = 8):
    """(F,H,W,3), (F,H,W,1) -> lists of full 8-frame clips; a trailing partial clip is zero padded like the reference's
    loader does (evaluate_ucf101.py:70-93)."""
    F = video.shape[0]
    n = (F + depth - 1) // depth
    v = np.zeros((n * depth,) + video.shape[1:], dtype=np.float32)
    b = np.zeros((n * depth,) + bbox.shape[1:], dtype=np.float32)
    v[:F], b[:F] = video, bbox
    return v.reshape((n, depth) + video.shape[1:]), b.reshape((n, depth) + bbox.shape[1:])


@torch.no_grad()
def evaluate_videos(model: torch.nn.Module, videos: Iterable[Tuple[np.ndarray, np.ndarray, int]], n_classes: int,
                    clip_batch_size: int = 14, device=None):
    """videos: iterable of (video (F,224,224,3) in [0,1], bbox (F,224,224,1) in {0,1}, label).
    Returns dict(accuracy, iou_threshs, fmAP, vmAP, frame_ious, video_ious, n_tot_frames, n_vids)."""
    dev = device or next(model.parameters()).device
    model.eval()
    iou_threshs = np.linspace(0, 1, 21)
    frame_ious = np.zeros((n_classes, 21))
    video_ious = np.zeros((n_classes, 21))
    n_tot_frames = np.zeros((n_classes, 1))
    n_vids = np.zeros((n_classes, 1))
    n_correct = 0
    for video, bbox, label in videos:
        if float(np.sum(bbox)) == 0.0:
            continue                                               # 'Video has no bounding boxes'
        clips, boxes = video_to_clips(np.asarray(video), np.asarray(bbox))
        preds, counts = [], []
        for i in range(0, clips.shape[0], clip_batch_size):
            x = torch.from_numpy(clips[i:i + clip_batch_size]).permute(0, 4, 1, 2, 3).contiguous().to(dev, non_blocking=True)
            gt = torch.from_numpy(boxes[i:i + clip_batch_size]).permute(0, 4, 1, 2, 3).contiguous().to(dev, non_blocking=True)
            empty = torch.full((x.shape[0], 1), 500, dtype=torch.int64, device=dev)
            seg, pred, _ = model(x, empty, empty, 0, 0)
            counts.append(ops.frame_iou_counts(seg.contiguous(), gt))
            preds.append(pred)
        cnt = torch.cat(counts).cpu().numpy().astype(np.int64)     # (frames, 3): inter, union, gt
        fin_pred = int(torch.cat(preds).mean(0).argmax())
        n_correct += int(fin_pred == int(label))
        vid_inter = vid_union = 0
        for inter, union, ngt in cnt:
            if ngt == 0:
                continue
            n_tot_frames[label] += 1
            vid_inter += inter
            vid_union += union
            frame_ious[label] += (inter / union >= iou_threshs)
        n_vids[label] += 1
        video_ious[label] += (vid_inter / vid_union >= iou_threshs)
    with np.errstate(divide="ignore", invalid="ignore"):
        fAP = frame_ious / n_tot_frames
        vAP = video_ious / n_vids
    return dict(accuracy=n_correct / max(1.0, float(np.sum(n_vids))), iou_threshs=iou_threshs, fmAP=np.nanmean(fAP, axis=0),
                vmAP=np.nanmean(vAP, axis=0), frame_ious=frame_ious, video_ious=video_ious, n_tot_frames=n_tot_frames, n_vids=n_vids)
