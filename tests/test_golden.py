"""Golden fixtures (tests/golden/*.json, produced by the literal restatement of the Scala arithmetic): the C oracle must
reproduce them on the CPU (here); the CUDA path reproduces them in tests/test_gpu_zfullsize.py."""
import glob
import json
import os

import numpy as np
import pytest

from oracle.oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.json")))


def load(path):
    f = json.load(open(path))
    f["row_ptr"] = np.array(f["row_ptr"], np.int64); f["col"] = np.array(f["col"], np.int32)
    f["val"] = np.array(f["val"], np.float32); f["label"] = np.array(f["label"], np.int8)
    return f


def flat_draws(f):
    return np.array([i for st in f["draws"] for b in st for i in b], dtype=np.int32)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_c_oracle_reproduces_golden(path):
    f = load(path)
    orc = Oracle(f["row_ptr"], f["col"], f["val"], f["label"], f["dim"], f["lambda"])
    d = orc.dim_sparsity(f["n_train"])
    np.testing.assert_allclose(d, f["dim_sparsity_weight_space"], rtol=0, atol=0)            # quirk Q3, exact
    orc.set_dim_sparsity(d)
    w, losses = orc.sync_steps(np.zeros(f["dim"]), flat_draws(f), [f["B"]] * f["K"], f["lr"], n_steps=len(f["draws"]))
    np.testing.assert_allclose(losses, f["step_losses"], rtol=1e-13)
    np.testing.assert_allclose(w, f["final_weights"], rtol=1e-12, atol=1e-16)
    g, _ = orc.gradient(w, f["probe"])
    np.testing.assert_allclose(g, f["probe_gradient"], rtol=1e-12, atol=0)
    assert (g == 0).tolist() == (np.array(f["probe_gradient"]) == 0).tolist()
    np.testing.assert_array_equal(orc.forward(w, f["probe"]), f["probe_predictions"])
    n = len(f["label"])
    loss, acc = orc.loss_acc(w, begin=f["n_train"], n=n - f["n_train"])
    assert acc == f["test_accuracy"] and loss == pytest.approx(f["test_loss"], rel=1e-13)
    wa = orc.async_run(np.zeros(f["dim"]), np.array(f["async_samples"], np.int32), 1, f["lr"])
    np.testing.assert_allclose(wa, f["async_final_weights"], rtol=1e-12, atol=1e-16)
