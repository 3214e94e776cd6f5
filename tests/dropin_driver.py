"""A SMALL driver written for this repo's tests (not a copy of any reference file) that walks the same import surface and
call sequence as the reference's training script (main_ucf101.py:23-30 imports, :50-150 step, :155-223 train loop,
:226-278 validate, :389-419 criteria / Adam / ReduceLROnPlateau / ramp): used to prove on the GPU box -- where the
reference checkout does not exist -- that `python -m b200caps.launch <script>` runs a reference-style script end to end
on libb200caps.so, including the torch-1.7-era idioms the launcher shims (verbose= kwarg, np.int, CPU tensor indexed
by a CUDA index, nn.BCEWithLogitsLoss(size_average=True))."""
import argparse
import sys

import numpy as np
import torch
import torch.nn as nn
from torch import optim
from torch.utils.data import DataLoader
from tensorboardX import SummaryWriter

from datasets.ucf_dataloader import UCF101DataLoader
from models.capsules_ucf101 import CapsNet
from utils import ramp_ups
from utils.helpers import measure_pixelwise_gradient, measure_pixelwise_var_v2
from utils.losses import DiceLoss, SpreadLoss, weighted_mse_loss
from utils.metrics import IOU2, get_accuracy


def step(args, model, crit, lab, unl, epoch, wt_ramp):
    data = torch.cat([lab["data"], unl["data"]]).type(torch.cuda.FloatTensor)
    fl_data = torch.cat([lab["aug_data"], unl["aug_data"]]).type(torch.cuda.FloatTensor)
    action = torch.cat([lab["action"], unl["action"]]).cuda()
    seg = torch.cat([lab["loc_msk"], unl["loc_msk"]])                       # stays on the HOST like in the reference
    labels = torch.cat([lab["label_vid"], unl["label_vid"]]).cuda()
    idx = torch.where(labels == 1)[0]                                        # CUDA index
    out, act, _ = model(data, action, labels, epoch, args.thresh_epoch)
    flip_op, _, _ = model(fl_data, action, labels, epoch, args.thresh_epoch)
    seg_l = seg[idx].float().cuda()                                          # CPU tensor indexed by a CUDA index
    loc = crit["bce"](out[idx], seg_l) + crit["dice"](out[idx], seg_l)
    cls, _ = crit["cls"](act[idx], action[idx])
    flipped = torch.flip(flip_op, [4])
    l2 = weighted_mse_loss(flipped, out, torch.ones_like(out))
    if args.gv:
        cons = weighted_mse_loss(flipped, out, measure_pixelwise_gradient(out).type(torch.cuda.FloatTensor))
    else:
        v1 = measure_pixelwise_var_v2(out, torch.flip(flipped, [2]), frames_cnt=args.n_frames).type(torch.cuda.FloatTensor)
        v2 = measure_pixelwise_var_v2(torch.flip(out, [2]), flipped, frames_cnt=args.n_frames).type(torch.cuda.FloatTensor)
        cons = wt_ramp * (weighted_mse_loss(flipped, out, v1) + weighted_mse_loss(flipped, out, torch.flip(v2, [2]))) + (1 - wt_ramp) * l2
    return out, act, action, loc + cls + args.wt_cons * cons


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bs", type=int, default=2)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--n_frames", type=int, default=5)
    ap.add_argument("--thresh_epoch", type=int, default=11)
    ap.add_argument("--wt_cons", type=float, default=0.1)
    ap.add_argument("--gv", action="store_true")
    args = ap.parse_args()
    torch.manual_seed(47)
    sets = [UCF101DataLoader("train", [224, 224], file_id="train_annots_20_labeled.pkl"),
            UCF101DataLoader("train", [224, 224], file_id="train_annots_80_unlabeled.pkl"),
            UCF101DataLoader("validation", [224, 224], file_id="test_annots.pkl")]
    lab_loader, unl_loader, val_loader = [DataLoader(s, batch_size=max(1, args.bs // 2), num_workers=0, shuffle=False) for s in sets]
    model = CapsNet().cuda()
    crit = dict(cls=SpreadLoss(num_class=24, m_min=0.2, m_max=0.9), bce=nn.BCEWithLogitsLoss(size_average=True), dice=DiceLoss())
    optimizer = optim.Adam(model.parameters(), lr=1e-4, weight_decay=0, eps=1e-6)
    scheduler = optim.lr_scheduler.ReduceLROnPlateau(optimizer, 'min', min_lr=1e-7, patience=5, factor=0.1, verbose=True)
    ramp = ramp_ups.exp_rampup(args.epochs)
    writer = SummaryWriter("unused")
    dummy = np.ones((2, 1), np.int) * 500
    for e in range(1, args.epochs + 1):
        model.train(mode=True)
        losses, accs = [], []
        lab_iter = iter(lab_loader)
        for unl in unl_loader:
            optimizer.zero_grad()
            out, act, action, loss = step(args, model, crit, next(lab_iter), unl, e, ramp(e))
            loss.backward()
            optimizer.step()
            losses.append(loss.item())
            accs.append(get_accuracy(act, action))
        model.eval()
        ious = []
        with torch.no_grad():
            for mb in val_loader:
                data = mb["data"].type(torch.cuda.FloatTensor)
                action = mb["action"].cuda()
                out, act, _ = model(data, action, torch.zeros(action.shape[0]).cuda(), 0, 0)
                mask = (out.cpu().numpy() > 0).astype(np.float32)
                for a in range(data.shape[0]):
                    ious.append(IOU2(mb["loc_msk"].numpy()[a], mask[a]))
        scheduler.step(float(np.mean(losses)))
        writer.add_scalars("train/loss", {"loss": float(np.mean(losses))}, e)
        assert all(np.isfinite(losses)), losses
        print(f"EPOCH DONE {e} loss {np.mean(losses):.4f} acc {np.mean(accs):.3f} iou {np.nanmean(ious):.3f} steps {len(losses)} dummy {int(dummy.sum())}")
    sys.stdout.flush()


if __name__ == "__main__":
    main()
