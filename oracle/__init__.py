"""CPU oracle for the nspeech spectrogram / Griffin-Lim hot path.

TEST INFRASTRUCTURE ONLY. Nothing under ``nspeech_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may, and there
only as the checker or the timed CPU baseline - never as the product path.

Parity status: the reference ships no tests or golden vectors, and its
arithmetic lives in un-vendored dependencies (librosa==0.6.0, scipy==1.0.0, numpy==1.14.2,
``requirements.txt:3,6,11``) that are absent here. So:
* ``oracle/librosa060.py`` restates librosa 0.6.0's published stft / istft / filters.mel algorithms;
  it is cross-checked against independent implementations (torch.stft / torch.istft /
  torchaudio.melscale_fbanks / analytic known answers) in ``tests/test_oracle.py``;
* the reference's OWN composition code (``neural_speech/utils/audio.py``, unmodified, imported from
  /root/reference with the absent third-party modules shimmed by that restatement) generated the
  committed fixtures under ``tests/golden/`` via ``tests/golden/make_golden.py``; ``oracle/audio_oracle.py``
  is pinned against those.
The dependency layer itself therefore remains "parity unpinned" against a real librosa 0.6.0 run.
"""
