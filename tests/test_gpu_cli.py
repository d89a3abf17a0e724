"""Drop-in surface on the GPU: `UrsoNet` facade + pose_estimator CLI on a synthetic URSO-format dataset
(BASELINE.json configs[0]: train --backbone resnet18 --image_scale 0.25 --batch_size 1, 2 steps on 4 frames),
checkpoint / resume naming, detect() parity against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import ursonet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    from ursonet_b200 import data as D
    root = tmp_path_factory.mktemp("urso")
    D.write_synthetic_urso(str(root / "datasets" / "synth"), n_train=4, n_val=2, n_test=2)
    return root


def run_cli(workdir, *argv):
    from ursonet_b200 import pose_estimator as PE
    return PE.main(list(argv) + ["--dataset", "synth", "--data_dir", str(workdir / "datasets"), "--logs", str(workdir / "logs"),
                                 "--backbone", "resnet18", "--image_scale", "0.25", "--ori_resolution", "8"])


def test_config1_train_then_resume_then_evaluate(workdir, capsys):
    model = run_cli(workdir, "train", "--weights", "none", "--batch_size", "1", "--epochs", "1", "--steps_per_epoch", "2")
    assert model.epoch == 1 and tuple(model.config.IMAGE_SHAPE) == (256, 320, 3)
    ck = model.checkpoint_path.format(epoch=1)
    assert os.path.exists(ck) and os.path.basename(ck) == "weights_synth_0001.npz"
    assert os.path.exists(os.path.join(model.log_dir, "config_0.json"))
    with np.load(ck) as z:
        assert "stage1_unit1_conv1/kernel" in z.files and z["conv0/kernel"].shape == (7, 7, 3, 64)
    w1 = model.engine.params.state_dict()
    # resume: --weights last picks the checkpoint, parses epoch 1 from its name and continues to epoch 2
    model2 = run_cli(workdir, "train", "--weights", "last", "--batch_size", "1", "--epochs", "2", "--steps_per_epoch", "2")
    assert model2.epoch == 2 and os.path.exists(model2.checkpoint_path.format(epoch=2))
    assert model2.log_dir == model.log_dir
    w2 = model2.engine.params.state_dict()
    assert not np.array_equal(w1["loc_final/kernel"], w2["loc_final/kernel"])
    # evaluate: loads the last checkpoint, batch 1 inference, writes the three CSVs of the reference
    model3 = run_cli(workdir, "evaluate", "--weights", "last")
    out = capsys.readouterr().out
    assert "Mean est. location error" in out and "ESA score" in out
    for f in ("ori_err.csv", "loc_err.csv", "dists_err.csv"):
        assert os.path.exists(os.path.join(str(workdir / "logs"), f))
    assert model3.mode == "inference"


def test_keras_h5_checkpoints(workdir):
    """The reference's checkpoint format (Keras HDF5 weight files, net.py:964-967,1120) through ursonet_b200/hdf5.py:
    save_weights('.h5') -> load_weights by name with an exclude list (net.py:816-852) -> identical tensors; `--weights
    <file>.h5` resumes from it with the epoch parsed from the name (net.py:956)."""
    from ursonet_b200 import hdf5
    model = run_cli(workdir, "train", "--weights", "none", "--batch_size", "1", "--epochs", "1", "--steps_per_epoch", "1")
    sd = model.engine.params.state_dict()
    h5 = os.path.join(model.log_dir, "weights_synth_0007.h5")
    model.save_weights(h5)
    with hdf5.File(h5, strict=True) as f:
        assert [bytes(x).decode() for x in f.attrs["layer_names"]][0] == "conv0"
        assert f["conv0/conv0/kernel:0"].shape == (7, 7, 3, 64)
    other = run_cli(workdir, "train", "--weights", h5, "--batch_size", "1", "--epochs", "7", "--steps_per_epoch", "1")
    assert other.epoch == 7 and other.log_dir == model.log_dir       # nothing left to train: weights are the file's
    got = other.engine.params.state_dict()
    assert set(got) == set(sd)
    for k in sd:
        assert np.array_equal(got[k], sd[k]), k
    # by-name load with an exclude pattern keeps the excluded layers' own initialisation
    from ursonet_b200 import net
    fresh = net.UrsoNet("training", model.config, str(workdir / "logs"))
    before = fresh.engine.params.state_dict()
    fresh.load_weights(h5, h5, by_name=True, exclude=["ori_final", "loc_final"])
    after = fresh.engine.params.state_dict()
    assert np.array_equal(after["ori_final/kernel"], before["ori_final/kernel"])
    assert np.array_equal(after["conv0/kernel"], sd["conv0/kernel"])
    # CHECKPOINT_FORMAT = 'h5' makes train() write Keras files
    model.config.CHECKPOINT_FORMAT = "h5"
    model.set_log_dir()
    assert model.checkpoint_path.endswith("weights_synth_{epoch:04d}.h5")


def test_train_with_device_sim2real(workdir):
    """--sim2real (BASELINE configs[3] turns it on): the uploaded uint8 frames go through the augmentation kernel when
    they are swapped in; what the network then sees is grey (3 equal channels) and matches the oracle for the drawn
    parameters (net.py:390-406)."""
    from oracle import sim2real_oracle as S
    model = run_cli(workdir, "train", "--weights", "none", "--batch_size", "2", "--epochs", "1", "--steps_per_epoch", "2",
                    "--sim2real")
    assert model.config.SIM2REAL_AUG and model.epoch == 1
    e = model.engine
    torch.cuda.synchronize()
    got = e.img_u8.cpu().numpy()
    assert np.array_equal(got[..., 0], got[..., 1]) and np.array_equal(got[..., 1], got[..., 2])
    # the batch now in the network's input buffer is the LAST one swapped in: staging buffer + its parameter records
    src = e._st[0].cpu().numpy()
    assert np.array_equal(got, S.augment_batch(src, model._aug))
    assert np.isfinite(e.losses.cpu().numpy()).all()


def test_detect_matches_oracle_and_asserts_batch_size(workdir):
    from ursonet_b200 import data as D, net, pose_estimator as PE
    args = PE.build_parser().parse_args(["evaluate", "--dataset", "synth", "--weights", "none", "--backbone", "resnet18",
                                         "--image_scale", "0.25", "--ori_resolution", "8", "--regress_ori"])
    cfg = PE.make_config(args)
    model = net.UrsoNet("inference", cfg, str(workdir / "logs"))
    p64 = O.init_weights(cfg, seed=5, pretrained_like=True)
    model.engine.params.load_state_dict({k: v.numpy() for k, v in p64.items()})
    ds = D.Urso(); ds.load_dataset(str(workdir / "datasets" / "synth"), cfg, "test")
    image = ds.load_image(0)
    res = model.detect([image])
    molded, _, _ = model.mold_inputs([image])
    rloc, rori = O.forward(p64, torch.from_numpy(molded).double(), cfg)
    assert np.allclose(res[0]["loc"], rloc[0].numpy(), rtol=5e-2, atol=5e-2 * np.abs(rloc.numpy()).max())
    assert abs(np.linalg.norm(res[0]["ori"]) - 1) < 1e-4                      # normalised quaternion output
    assert abs(float(np.dot(res[0]["ori"], rori[0].numpy()))) > 0.995
    with pytest.raises(AssertionError, match="BATCH_SIZE"):
        model.detect([image, image])
    with pytest.raises(AssertionError, match="inference"):
        net.UrsoNet.detect(type("M", (), {"mode": "training", "config": cfg})(), [image])


def test_missing_extension_fails_loudly(monkeypatch):
    from ursonet_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/liburso_b200.so")
    with pytest.raises(lib.UrsoError, match="no CPU / eager fallback"):
        lib.load()


def test_f16_flag_runs_on_the_16_bit_engine(workdir):
    """--f16 (net.py:590-593) is served by the bf16-storage engine with the reference's fp16 Adam epsilon."""
    model = run_cli(workdir, "train", "--weights", "none", "--batch_size", "1", "--epochs", "1", "--steps_per_epoch", "1",
                    "--f16")
    assert model.config.F16 and model.epoch == 1
    assert torch.isfinite(model.engine.params.flat).all()


def test_ctypes_example_of_integration_md_runs():
    """examples/conv2d_ctypes.py: a stand-alone ctypes binding of one Conv2D operator (no ursonet_b200 import)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "conv2d_ctypes.py")], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "max rel err" in r.stdout
