"""ORACLE — TEST INFRASTRUCTURE ONLY.

A CPU restatement (plain ``torch`` ops, fp32 or fp64, no custom kernels) of the
reference algorithm for the one hot path this repository accelerates: the
wenet-e2e/wesep target-speaker-extraction *train step* (Spex+/ConvTasNet forward,
SI-SDR + CE loss, backward via torch autograd of the restated forward, per-tensor
gradient clip, Adam).  Every function cites the reference file:line it follows.

``oracle/bsrnn.py`` restates the pBSRNN forward (STFT/iSTFT, band split, BLSTM
recurrence, mask head) the same way; it is pinned by ``tests/golden/bsrnn_*.npz`` (``make_golden_bsrnn.py``) and
is groundwork for the next round's CUDA path — nothing in ``wesep_b200`` uses it.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package — as the checker or the timed CPU
baseline, never as part of the product path.  ``wesep_b200`` must not import it.

Parity pinning: the reference ships NO tests, golden vectors or fixtures for this path
 and its SI-SDR arithmetic lives in the un-pinned third-party package
``auraloss`` (wesep/utils/losses.py:24-25; requirements.txt:26, absent here), restated
in ``oracle/losses.py`` from its published formula and cross-checked with the in-tree
numpy ``cal_SISNR`` (wesep/utils/score.py:7-21).  The restatement is instead pinned
against OUTPUTS OF THE REFERENCE ITSELF run in the build container:
``tests/golden/make_golden.py`` imports the real ``wesep`` modules from /root/reference
(through ``oracle/stubs``) and stores their outputs/gradients as fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this package against them.
"""
