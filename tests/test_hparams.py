import os

import pytest

from nspeech_b200 import hparams


def test_load_defaults_match_reference_yaml():
    hp = hparams.load()
    assert hp.num_freq == 1025 and hp.num_mels == 80 and hp.sample_rate == 20000
    assert hp.frame_shift_ms == 12.5 and hp.frame_length_ms == 50
    assert hp.preemphasis == 0.97 and hp.ref_level_db == 20 and hp.min_level_db == 100
    assert hp.power == 1.5 and hp.griffin_lim_iters == 60 and hp.max_iters == 300
    assert hparams.get_hparams() is hp


def test_parse_overrides_and_types():
    hp = hparams.load()
    hp.parse("griffin_lim_iters=100,power=1.2,min_level_db=-100")
    assert hp.griffin_lim_iters == 100 and isinstance(hp.griffin_lim_iters, int)
    assert hp.power == 1.2 and hp.min_level_db == -100
    with pytest.raises(ValueError):
        hp.parse("no_such_key=1")
    k1 = hp.audio_key()
    hp.parse("sample_rate=22050")
    assert hp.audio_key() != k1
    hparams.load()


def test_merge_order_and_model_yaml(tmp_path):
    (tmp_path / "audio.yaml").write_text("num_freq: 513\nsample_rate: 16000\nx: 1\n")
    (tmp_path / "train.yaml").write_text("x: 2\nbatch_size: 32\n")
    (tmp_path / "taco1.yaml").write_text("x: 3\noutputs_per_step: 5\n")
    hp = hparams.load("taco1", path=str(tmp_path))
    assert hp.x == 3 and hp.batch_size == 32 and hp.outputs_per_step == 5 and hp.num_freq == 513
    assert "outputs_per_step" in hparams.debug_string(hp)
    hparams.load()


def test_get_before_load_raises():
    old = hparams._hparams
    hparams._hparams = None
    try:
        with pytest.raises(RuntimeError):
            hparams.get_hparams()
    finally:
        hparams._hparams = old
