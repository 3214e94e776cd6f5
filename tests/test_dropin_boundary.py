"""The drop-in boundary: with this repo's package directory on PYTHONPATH the names
a reference script imports -- models.*, utils.losses / helpers / ramp_ups / metrics, datasets.* -- must resolve INTO
THIS REPO even when the script directory (first on sys.path) holds the reference's own namespace directories, and
the compat launcher's shims must make the scripts' torch-1.7-era calls work.  CPU only."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pi-consistency-activity-detection_b200")
REF = "/root/reference"


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = PKG
    return env


@pytest.fixture()
def fake_reference_tree(tmp_path):
    """A directory shaped like the reference checkout: models/ utils/ datasets/ WITHOUT __init__.py (namespace
    directories), each module marked so a wrong resolution is visible, plus a script with the reference's imports."""
    for d, files in (("models", ("capsules_ucf101", "pytorch_i3d")), ("utils", ("losses", "helpers", "ramp_ups", "metrics")),
                     ("datasets", ("ucf_dataloader",))):
        os.makedirs(tmp_path / d)
        for f in files:
            (tmp_path / d / f"{f}.py").write_text("ORIGIN = 'fake-reference'\nraise RuntimeError('reference module imported')\n")
    (tmp_path / "main_fake.py").write_text(textwrap.dedent("""
        import numpy as np
        import torch
        import imageio
        from torch import optim
        from tensorboardX import SummaryWriter
        from datasets.ucf_dataloader import UCF101DataLoader
        from models.capsules_ucf101 import CapsNet
        from utils.losses import SpreadLoss, DiceLoss, weighted_mse_loss
        from utils.metrics import get_accuracy, IOU2
        from utils.helpers import measure_pixelwise_var_v2, measure_pixelwise_gradient
        from utils import ramp_ups

        def torch17_era_calls():
            lin = torch.nn.Linear(2, 2)
            opt = optim.Adam(lin.parameters(), lr=1e-3, weight_decay=0, eps=1e-6)
            sch = optim.lr_scheduler.ReduceLROnPlateau(opt, 'min', min_lr=1e-7, patience=5, factor=0.1, verbose=True)
            sch.step(1.0)
            z = np.ones((3, 1), np.int) * 500
            crit = torch.nn.BCEWithLogitsLoss(size_average=True)
            l = crit(torch.zeros(2, 3), torch.ones(2, 3))
            SummaryWriter('x').add_scalars('a', {'b': 1.0}, 1)
            return float(l), int(z.sum()), ramp_ups.exp_rampup(100)(1)

        if __name__ == '__main__':
            print('RESULT', torch17_era_calls())
    """))
    return tmp_path


def test_plain_pythonpath_recipe_resolves_into_repo(fake_reference_tree):
    """INTEGRATION.md section 2 without the launcher: cwd = reference tree, PYTHONPATH = this package."""
    code = ("import importlib.util as u, json; "
            "print(json.dumps({n: u.find_spec(n).origin for n in ('utils.losses','utils.helpers','utils.ramp_ups','utils.metrics',"
            "'models.capsules_ucf101','models.pytorch_i3d','datasets.ucf_dataloader','datasets.load_jhmdb_pytorch_multi')}))")
    r = subprocess.run([sys.executable, "-c", code], cwd=fake_reference_tree, env=_env(), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    for name, origin in rep.items():
        assert origin and os.path.abspath(origin).startswith(PKG), (name, origin)


def test_launcher_check_on_fake_tree(fake_reference_tree):
    r = subprocess.run([sys.executable, "-m", "b200caps.launch", "--check", "main_fake.py"], cwd=fake_reference_tree, env=_env(),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.loads(r.stdout[r.stdout.index("{"):])
    for name, origin in rep["resolves"].items():
        assert os.path.abspath(origin).startswith(PKG), (name, origin)
    for name in ("UCF101DataLoader", "CapsNet", "SpreadLoss", "DiceLoss", "weighted_mse_loss", "get_accuracy", "IOU2",
                 "measure_pixelwise_var_v2", "measure_pixelwise_gradient"):
        assert os.path.abspath(rep["names_bound_by_script"][name]).startswith(PKG), name


def test_launcher_runs_script_with_torch17_era_calls(fake_reference_tree):
    """ReduceLROnPlateau(verbose=True), np.int, BCEWithLogitsLoss(size_average=True), tensorboardX stub."""
    r = subprocess.run([sys.executable, "-m", "b200caps.launch", "main_fake.py"], cwd=fake_reference_tree, env=_env(),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1]
    bce, zsum, ramp = eval(line[len("RESULT"):])
    assert abs(bce - float(np.log(2.0))) < 1e-6 and zsum == 1500
    assert abs(ramp - float(np.exp(-5.0 * 0.99 ** 2))) < 1e-12


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "main_ucf101.py")), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("script", ["main_ucf101.py", "main_jhmdb.py"])
def test_launcher_check_on_the_real_reference_scripts(script):
    """The UNMODIFIED reference scripts, imported under the launcher from the reference checkout: every hot-path name
    they bind comes from this repo (build container only -- /root/reference does not exist on the GPU box)."""
    r = subprocess.run([sys.executable, "-m", "b200caps.launch", "--check", script], cwd=REF, env=_env(), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    rep = json.loads(r.stdout[r.stdout.index("{"):])
    assert all(os.path.abspath(o).startswith(PKG) for o in rep["resolves"].values()), rep["resolves"]
    bound = rep["names_bound_by_script"]
    assert {"SpreadLoss", "DiceLoss", "weighted_mse_loss", "measure_pixelwise_var_v2", "measure_pixelwise_gradient"} <= set(bound)
    assert all(os.path.abspath(f).startswith(PKG) for f in bound.values()), bound


def test_synthetic_datasets_honour_the_reference_sample_format():
    """ucf_dataloader.py:189 / jhmdb_dataloader.py:229: dict keys, shapes, value ranges; deterministic per index."""
    sys.path.insert(0, PKG)
    try:
        from datasets.load_jhmdb_pytorch_multi import JHMDB
        from datasets.ucf_dataloader import UCF101DataLoader
        from datasets.ucf_dataloader_eval import UCF101DataLoader as UCFEval
    finally:
        sys.path.remove(PKG)
    lab = UCF101DataLoader("train", [224, 224], file_id="train_annots_20_labeled.pkl", use_random_start_frame=False)
    unl = UCF101DataLoader("train", [224, 224], file_id="train_annots_80_unlabeled.pkl", use_random_start_frame=False)
    val = UCF101DataLoader("validation", [224, 224], file_id="test_annots.pkl", use_random_start_frame=False)
    s = lab[3]
    assert set(s) == {"data", "loc_msk", "action", "aug_data", "label_vid"}
    assert tuple(s["data"].shape) == (3, 8, 224, 224) and tuple(s["loc_msk"].shape) == (1, 8, 224, 224)
    assert tuple(s["action"].shape) == (1,) and 0 <= int(s["action"]) < 24
    assert torch.equal(s["aug_data"], torch.flip(s["data"], [3]))
    assert float(s["data"].min()) >= 0 and float(s["data"].max()) < 1 and set(s["loc_msk"].unique().tolist()) <= {0.0, 1.0}
    assert s["label_vid"] == 1 and unl[0]["label_vid"] == 0 and val[0]["label_vid"] == 1
    assert torch.equal(lab[3]["data"], s["data"]) and not torch.equal(lab[4]["data"], s["data"])
    batch = next(iter(torch.utils.data.DataLoader(lab, batch_size=2, shuffle=False)))
    assert tuple(batch["data"].shape) == (2, 3, 8, 224, 224) and tuple(batch["action"].shape) == (2, 1)
    assert tuple(batch["label_vid"].shape) == (2,)
    j = JHMDB("train", [224, 224], file_id="jhmdb_classlist_33_33_labeled.txt", use_random_start_frame=False)[0]
    assert "label_vid" not in j and "mask_cls" in j and 0 <= int(j["action"]) < 21
    video, bbox, label = UCFEval("validation", [224, 224], 1, file_id="testing_annots.pkl", use_random_start_frame=False)[0]
    assert video.shape[1:] == (224, 224, 3) and bbox.shape[1:] == (224, 224, 1) and video.shape[0] == bbox.shape[0]


def test_metrics_and_ramps_known_answers():
    sys.path.insert(0, PKG)
    try:
        from utils import metrics, ramp_ups
    finally:
        sys.path.remove(PKG)
    pred = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1], [0.2, 0.3, 0.5]])
    assert metrics.get_accuracy(pred, torch.tensor([[1.0], [2.0], [2.0]])) == pytest.approx(2 / 3)
    gt = np.zeros((1, 2, 4, 4), dtype=np.float32)
    gt[..., :2, :] = 1
    out = np.zeros_like(gt)
    out[..., 1:3, :] = 1
    assert metrics.IOU2(gt, out) == pytest.approx(1 / 3)
    assert np.isnan(metrics.IOU2(np.zeros_like(gt), out))
    f = ramp_ups.exp_rampup(100)
    assert f(0) == pytest.approx(np.exp(-5.0)) and f(1) == pytest.approx(np.exp(-5.0 * 0.99 ** 2)) and f(100) == 1.0 and f(250) == 1.0
    assert ramp_ups.linear_rampup(10)(5) == pytest.approx(0.5) and ramp_ups.pseudo_rampup(5, 15)(10) == pytest.approx(0.5)


def test_no_product_module_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the shipped package may import it (a product path routed
    through the oracle would void every parity claim)."""
    bad = []
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                if "import oracle" in src or "from oracle" in src or "restate" in src:
                    bad.append(os.path.join(d, f))
    assert not bad, bad
