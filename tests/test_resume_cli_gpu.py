"""`train(train, restore)` round trips (/root/reference src/distilation/main.py:9-27, lstm_train.py:86-87,102-107,199):
the CLI writes the checkpoint `-r` and `-ch` read, the teacher weights reach the loops, and a restored LSTM loop continues bit-identically
(student + Adam moments, env state incl. the low parts of the joint angles, Dataset ring + sampling counter, carried acting state)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cli_train_restore_check_round_trip(tmp_path, monkeypatch, capsys):
    from reacherdistilation_b200 import config, main, mlp_train
    from reacherdistilation_b200.teacher import init_policy_params
    monkeypatch.setattr(config, "base_path", str(tmp_path))
    monkeypatch.setattr(mlp_train, "base_path", str(tmp_path))
    np.savez(tmp_path / "teacher.npz", params=init_policy_params(seed=3, final_std=0.3))      # picked up as <base_path>/teacher.npz
    a = main.main(["-ct", "--num_envs", "256", "--iterations", "40"])
    assert a["checkpoint"] == str(tmp_path / "student_mlp_b200.pt") and (tmp_path / "student_mlp_b200.pt").exists()
    assert "teacher.npz" in a["teacher"] and not a["resumed"]
    pa = a["trainer"].student.params.clone(); a["trainer"].close()
    b = main.main(["-ct", "-r", "--num_envs", "256", "--iterations", "0"])                     # restore only: the loop state comes back
    assert b["resumed"] and torch.equal(b["trainer"].student.params, pa) and b["trainer"].iteration == 40
    b["trainer"].close()
    shapes = main.main(["-ch"])
    assert shapes["student"]["params"] == (24380,) and shapes["iteration"] == 40 and shapes["env"]["qpos_lo"] == (256, 2)
    assert "checking saved variables" in capsys.readouterr().out


def test_no_teacher_weights_is_announced_and_restore_without_weights_raises(tmp_path, monkeypatch, capsys):
    from reacherdistilation_b200 import ReacherB200Error, config, mlp_train
    from reacherdistilation_b200.teacher import TeacherAgent
    monkeypatch.setattr(config, "base_path", str(tmp_path))
    monkeypatch.setattr(mlp_train, "base_path", str(tmp_path))
    out = mlp_train.train(True, False, num_envs=128, iterations=5, checkpoint=str(tmp_path / "c.pt"))
    assert out["teacher"].startswith("SYNTHETIC") and "SYNTHETIC" in capsys.readouterr().out
    out["trainer"].close()
    with pytest.raises(ReacherB200Error):
        TeacherAgent(restore=True)


def test_lstm_loop_resume_is_bit_exact(tmp_path):
    from reacherdistilation_b200 import lstm_train
    kw = dict(num_envs=8, generations=8, verbose=False, lr=1e-3)
    a = lstm_train.train(True, False, iterations=100, checkpoint=str(tmp_path / "a.pt"), **kw)
    b = lstm_train.train(True, False, iterations=50, checkpoint=str(tmp_path / "b.pt"), **kw)
    for k in ("env", "dataset"):
        b[k].close()
    c = lstm_train.train(True, True, iterations=100, checkpoint=str(tmp_path / "b.pt"), **kw)
    assert c["resumed"] and c["iterations"] == 100
    assert torch.equal(a["student"].params, c["student"].params) and torch.equal(a["student"].m, c["student"].m)
    assert a["losses"] == c["losses"] and a["rewards"] == c["rewards"]
    sa, sc = a["env"].get_state(), c["env"].get_state()
    for k in sa:
        assert torch.equal(sa[k], sc[k]), k
    da, dc = a["dataset"].state_dict(), c["dataset"].state_dict()
    assert da["generations"] == dc["generations"] and da["draw"] == dc["draw"] and torch.equal(da["ob"], dc["ob"]) and torch.equal(da["s"], dc["s"])
    for r in (a, c):
        r["env"].close(); r["dataset"].close()
