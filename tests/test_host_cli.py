"""Host-side mirror of the reference's entry points (no GPU): the constants of `config.py:15-32`, the CLI flags of `main.py:9-27` and how
they reach `mlp_train.train(train, restore)` / `lstm_train.train(train, restore)`."""
import pytest


def test_config_constants_match_the_reference_values():
    from reacherdistilation_b200 import config as c
    # /root/reference src/distilation/config.py:15-32
    assert (c.EPISODE_STEPS, c.OBSPACE_SHAPE, c.ACSPACE_SHAPE, c.PDFLAT_SHAPE, c.GAMMA) == (50, 11, 2, 4, 0.99)
    assert (c.TOTAL_EPISODES, c.STEPS_UNROLLED, c.LSTM_BATCH_SIZE, c.MLP_BATCH_SIZE, c.NUM_UNITS) == (8000, 10, 20, 20, 200)
    assert (c.KEEP_PROB, c.MAX_CAPACITY, c.TRAINING_EPOCHS) == (0.5, 10, 1)


@pytest.mark.parametrize("argv,which,restore,kp", [
    (["-ct"], "mlp", False, 0.5),
    (["--mlp_train", "-r"], "mlp", True, 0.5),
    (["-lt", "-k", "0.8"], "lstm", False, 0.8),
    (["--lstm_train", "--restore", "--keep_prob", "1.0"], "lstm", True, 1.0),
    (["-lt", "-ct"], "lstm", False, 0.5),               # main.py:24-27: `elif` chain, -lt wins over -ct
])
def test_cli_flags_dispatch_like_the_reference(monkeypatch, argv, which, restore, kp):
    from reacherdistilation_b200 import lstm_train, main, mlp_train
    calls = []
    monkeypatch.setattr(mlp_train, "train", lambda train, restore, **kw: calls.append(("mlp", train, restore, kw)))
    monkeypatch.setattr(lstm_train, "train", lambda train, restore, **kw: calls.append(("lstm", train, restore, kw)))
    main.main(argv + ["--num_envs", "8", "--iterations", "3"])
    assert len(calls) == 1
    name, train, rest, kw = calls[0]
    assert name == which and train is True and rest is restore
    assert kw["keep_prob"] == pytest.approx(kp) and kw["num_envs"] == 8 and kw["iterations"] == 3


def test_cli_without_a_mode_flag_does_nothing(monkeypatch):
    from reacherdistilation_b200 import lstm_train, main, mlp_train
    monkeypatch.setattr(mlp_train, "train", lambda *a, **k: pytest.fail("must not train"))
    monkeypatch.setattr(lstm_train, "train", lambda *a, **k: pytest.fail("must not train"))
    main.main([])                                        # main.py:21-27 falls through when no flag is given


def test_cli_two_headed_flag_reaches_the_backup_loop(monkeypatch):
    """`-lt --two_headed`: lstm_train(train, keep, lstm_trained_data_path, restore) of backup/student_rollout.py:262,791-792."""
    from reacherdistilation_b200 import lstm2_train, lstm_train, main
    calls = []
    monkeypatch.setattr(lstm_train, "train", lambda *a, **k: pytest.fail("the two-headed loop was asked for"))
    monkeypatch.setattr(lstm2_train, "lstm_train", lambda train, keep, path, restore, **kw: calls.append((train, keep, path, restore, kw)))
    main.main(["-lt", "--two_headed", "-k", "0.7", "-r", "--num_envs", "8", "--iterations", "3", "--checkpoint", "/tmp/x.pt"])
    assert len(calls) == 1
    train, keep, path, restore, kw = calls[0]
    assert train is True and keep == pytest.approx(0.7) and path == "/tmp/x.pt" and restore is True and kw["num_envs"] == 8 and kw["iterations"] == 3


def test_lstm2_spec_defaults_are_the_checked_in_source():
    """backup/student_rollout.py:48-50 (NUM_UNITS 1, STEPS_UNROLLED 2), :156 (state never reassigned), :158-172 (trunk 128, reward 64 -> 1, action 64 -> 4)."""
    import numpy as np
    from oracle import lstm2_np as L2
    from reacherdistilation_b200.student_nn import LSTM2_TFEVENTS_SPEC, init_lstm2_params, lstm2_spec
    assert tuple(int(v) for v in lstm2_spec()[:7]) == L2.SOURCE_SPEC(1, 2)
    assert tuple(int(v) for v in lstm2_spec(**LSTM2_TFEVENTS_SPEC)[:9]) == L2.TFEVENTS_SPEC
    for spec in (L2.SOURCE_SPEC(100, 20), L2.TFEVENTS_SPEC):
        arr = L2.Layout(spec).spec_array()
        assert np.array_equal(init_lstm2_params(arr, 3), L2.init_params(spec, 3)) and init_lstm2_params(arr, 3).size == L2.param_count(spec)
