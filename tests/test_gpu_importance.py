"""B200ImportanceFlowModel: the multi-flow surface of
/root/reference/src/nessai/flowmodel/importance.py (dtype / shape contract of
tests/test_flowmodel/test_flowmodel_importance.py) on the kernels."""

import os
import pickle

import numpy as np
import pytest
import torch
from conftest import load_golden

pytestmark = pytest.mark.gpu


def make(tmp_path, name="c2_realnvp_resnet", **training):
    from nessai_b200.importance import B200ImportanceFlowModel

    g, cfg, sd = load_golden(name)
    tc = dict(device_tag="cuda:0")
    tc.update(training)
    ifm = B200ImportanceFlowModel(flow_config=cfg, training_config=tc, output=str(tmp_path))
    ifm.initialise()
    return ifm, g, cfg, sd


def test_log_prob_all_matches_each_flow_and_oracle(tmp_path):
    from test_oracle import numpy_flow

    ifm, g, cfg, sd = make(tmp_path)
    assert ifm.model is None and ifm.n_models == 0
    rng = np.random.default_rng(0)
    sds = []
    for k in range(3):
        ifm.add_new_flow(reset=(k != 1))
        assert ifm.n_models == k + 1
        sdk = {key: (np.asarray(v) + (0.02 * k * rng.standard_normal(np.shape(v))).astype(np.float32)
                     if np.asarray(v).dtype.kind == "f" and "running_var" not in key else np.asarray(v))
               for key, v in sd.items()}
        ifm.model.load_state_dict({key: torch.from_numpy(np.asarray(v)) for key, v in sdk.items()})
        sds.append(sdk)
    x = np.asarray(g["x"], dtype=np.float64)
    lp = ifm.log_prob_all(x)
    assert lp.shape == (len(x), 3) and lp.dtype == np.float64
    for k in range(3):
        lk = ifm.log_prob_ith(x, k)
        assert lk.dtype == np.float64
        np.testing.assert_array_equal(lp[:, k], lk)
        ref = numpy_flow(cfg, sds[k]).log_prob(x)
        np.testing.assert_allclose(lk, ref, rtol=1e-4, atol=1e-4)
    s = ifm.sample_ith(1, N=100)
    assert s.shape == (100, cfg["n_inputs"]) and s.dtype == np.float64


def test_train_save_and_reload_all(tmp_path):
    ifm, g, cfg, sd = make(tmp_path, "c2_realnvp_mlp", max_epochs=3, patience=3)
    data = np.asarray(g["train_data"])
    for level in range(2):
        ifm.add_new_flow(reset=True)
        out = os.path.join(str(tmp_path), f"level_{level}")
        hist = ifm.train(data, output=out, plot=False)
        assert len(hist["loss"]) == 3 and np.isfinite(hist["loss"]).all()
    assert len(ifm.weights_files) == 2
    x = np.asarray(g["x"], dtype=np.float64)
    before = ifm.log_prob_all(x)
    state = pickle.loads(pickle.dumps(ifm))
    assert state.models is None and state._resume_n_models == 2 and state.initialised is False
    state.resume(cfg, weights_path=str(tmp_path))
    assert state.n_models == 2
    np.testing.assert_allclose(state.log_prob_all(x), before, rtol=1e-6, atol=1e-6)


def test_levels_reuse_the_trainer_and_match_fresh_trainers(tmp_path):
    """Every level of the importance sampler adds a flow of the same architecture
    (flowmodel/importance.py:80-99) and trains it: the trainer (plan, workspaces) is reused across
    the levels with fresh optimiser state, and gives exactly what a fresh trainer per level gives."""
    from nessai_b200.importance import B200ImportanceFlowModel

    g, cfg, sd = load_golden("c2_realnvp_resnet")
    data = np.asarray(g["train_data"])
    rng = np.random.default_rng(3)
    sets = [data[rng.permutation(len(data))[:1500]] * s for s in (1.0, 1.3, 0.8)]
    out = {}
    for reuse in (True, False):
        torch.manual_seed(11)
        ifm = B200ImportanceFlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0", max_epochs=4, patience=4),
                                      output=str(tmp_path / str(reuse)), rng=np.random.default_rng(5))
        ifm.initialise()
        trainers, hists = [], []
        for level, xs in enumerate(sets):
            ifm.add_new_flow(reset=True)
            if not reuse:
                ifm._fused = None  # a fresh trainer per level
            hists.append(ifm.train(xs, output=os.path.join(ifm.output, f"level_{level}"), plot=False))
            trainers.append(ifm._fused)
        if reuse:
            assert trainers[0] is trainers[1] is trainers[2]
        out[reuse] = (hists, ifm.log_prob_all(np.asarray(g["x"], dtype=np.float64)))
    for ha, hb in zip(out[True][0], out[False][0]):
        np.testing.assert_array_equal(ha["loss"], hb["loss"])
        np.testing.assert_array_equal(ha["val_loss"], hb["val_loss"])
    np.testing.assert_array_equal(out[True][1], out[False][1])
    lp = out[True][1]
    assert lp.shape[1] == 3 and np.isfinite(lp).all() and not np.allclose(lp[:, 0], lp[:, 1])
