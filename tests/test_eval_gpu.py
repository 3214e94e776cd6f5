"""-m gpu: the evaluation consumer (evaluate_ucf101.py:95-187): eval forward with BatchNorm folded into
the convolutions at the reference's clip batch of 14, device-side per-frame IoU counts, f-mAP / v-mAP bookkeeping."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model():
    from models.capsules_ucf101 import CapsNet
    from oracle import restate
    sd = restate.make_state_dict(24, seed=0)
    # non-trivial running statistics, as after training
    g = torch.Generator().manual_seed(3)
    for k in sd:
        if k.endswith("running_mean"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.75
    m = CapsNet(pt_path=None)
    m.load_state_dict(sd)
    return m.cuda().eval(), sd


def test_folded_batchnorm_eval_forward_batch14():
    """conv + folded BN + ReLU in one kernel == conv -> running-statistics BatchNorm kernel -> ReLU, at 14 clips; and both
    against the oracle restatement (1 clip, fp64)."""
    from b200caps import engine
    from oracle import restate
    model, sd = _model()
    g = torch.Generator().manual_seed(14)
    x = torch.rand((14, 3, 8, 224, 224), generator=g).cuda()
    empty = torch.full((14, 1), 500, dtype=torch.int64).cuda()
    with torch.no_grad():
        assert engine.EVAL_FOLD_BN
        seg_f, act_f, _ = model(x, empty, empty, 0, 0)
        engine.EVAL_FOLD_BN = False
        try:
            seg_u, act_u, _ = model(x, empty, empty, 0, 0)
        finally:
            engine.EVAL_FOLD_BN = True
    torch.cuda.synchronize()
    da = float((act_f - act_u).abs().max() / act_u.abs().max())
    same_cls = act_f.argmax(1).tolist() == act_u.argmax(1).tolist()
    print(f"folded vs unfolded eval (14 clips, bf16): act {da:.2e}, same classes {same_cls}")
    assert da < 2e-2
    # oracle, clip 0 (the eval forward is per-clip independent: running statistics, no dropout)
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        o_ref, a_ref, _ = restate.capsnet_forward(sd64, x[:1].double().cpu(), empty[:1].double().cpu(), empty[:1].double().cpu(),
                                                  0, 0, False, None, None, 24)
    e_act = float((act_f[:1].double().cpu() - a_ref).abs().max() / a_ref.abs().max())
    print(f"folded eval vs fp64 oracle (clip 0): act {e_act:.2e}")
    assert e_act < 2e-2
    top2 = a_ref.topk(2).values[0]
    if int(act_f[0].argmax()) == int(a_ref.argmax(1)):
        l2 = float((seg_f[:1].double().cpu() - o_ref).norm() / o_ref.norm())
        print(f"   logits rel-L2 {l2:.2e}")
        assert l2 < 3e-2
    else:
        assert float(top2[0] - top2[1]) / float(a_ref.abs().max()) < 2e-2      # a tie the bf16 path may break either way


def test_device_iou_counts_and_map_bookkeeping():
    """evaluate_videos (device-side counts) against a numpy restatement of evaluate_ucf101.py:127-187 fed with the SAME
    model outputs copied to the host the way the reference does."""
    from b200caps import ops
    from b200caps.evaluate import evaluate_videos, video_to_clips
    from datasets.ucf_dataloader_eval import UCF101DataLoader
    model, _ = _model()
    os.environ.setdefault("B200CAPS_SYNTH_LEN", "2,2,3")
    ds = UCF101DataLoader("validation", [224, 224], 1, file_id="test.txt", use_random_start_frame=False)
    vids = [ds[i] for i in range(min(3, len(ds)))]
    # kernel vs numpy on random logits / masks (includes logits within 1e-8 of zero: sigmoid rounds to exactly 0.5)
    g = torch.Generator().manual_seed(1)
    lg = torch.randn((5, 1, 8, 224, 224), generator=g)
    lg.view(-1)[:1000] = torch.linspace(-1e-7, 1e-7, 1000)
    gt = (torch.rand((5, 1, 8, 224, 224), generator=g) > 0.7).float()
    cnt = ops.frame_iou_counts(lg.cuda(), gt.cuda()).cpu().numpy()
    p = (torch.sigmoid(lg) >= 0.5).numpy().reshape(40, -1)
    t = (gt > 0).numpy().reshape(40, -1)
    ref = np.stack([(p & t).sum(1), (p | t).sum(1), t.sum(1)], 1)
    assert np.array_equal(cnt, ref)
    res = evaluate_videos(model, vids, 24, clip_batch_size=14)
    # numpy restatement on host copies of the model's outputs
    iou_threshs = np.linspace(0, 1, 21)
    frame_ious, video_ious = np.zeros((24, 21)), np.zeros((24, 21))
    n_tot, n_vids, n_correct = np.zeros((24, 1)), np.zeros((24, 1)), 0
    with torch.no_grad():
        for video, bbox, label in vids:
            clips, boxes = video_to_clips(np.asarray(video), np.asarray(bbox))
            x = torch.from_numpy(clips).permute(0, 4, 1, 2, 3).contiguous().cuda()
            empty = torch.full((x.shape[0], 1), 500, dtype=torch.int64).cuda()
            seg, pred, _ = model(x, empty, empty, 0, 0)
            seg_np = np.transpose(torch.sigmoid(seg).cpu().numpy(), [0, 2, 3, 4, 1]).reshape((-1, 224, 224, 1))
            gt_np = boxes.reshape((-1, 224, 224, 1))
            n_correct += int(np.argmax(np.mean(pred.cpu().numpy(), axis=0)) == label)
            spg = (seg_np >= 0.5).astype(np.int64) + gt_np
            vi = vu = 0
            for i in range(gt_np.shape[0]):
                if np.sum(gt_np[i]) == 0:
                    continue
                n_tot[label] += 1
                inter, union = np.count_nonzero(spg[i] == 2), np.count_nonzero(spg[i])
                vi, vu = vi + inter, vu + union
                frame_ious[label] += (inter / union >= iou_threshs)
            n_vids[label] += 1
            video_ious[label] += (vi / vu >= iou_threshs)
    assert np.array_equal(res["frame_ious"], frame_ious) and np.array_equal(res["video_ious"], video_ious)
    assert np.array_equal(res["n_tot_frames"], n_tot) and np.array_equal(res["n_vids"], n_vids)
    assert abs(res["accuracy"] - n_correct / float(np.sum(n_vids))) < 1e-12
    print("eval consumer:", {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in res.items() if k in ("accuracy",)},
          "fmAP@0.5", res["fmAP"][10], "vmAP@0.5", res["vmAP"][10])
