"""CPU oracle for the ffsubsync alignment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The product package ``ffsubsync_b200`` never imports it
and raises when its CUDA library is missing.

Each function is a numpy/float64 restatement of a piece of the reference
(``/root/reference/ffsubsync``) and cites the file:line it follows.  Parity
status per module:

* ``aligner_oracle``  - PINNED against the reference's own ``aligners.py`` run
  in the build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.json|npz``)
  and against the KATs of ``tests/test_alignment.py`` / ``tests/test_multi_segment.py``.
* ``raster_oracle``   - PINNED against ``SubtitleScaler`` + ``SubtitleSpeechTransformer``
  run in the build container, and the ``tests/test_subtitles.py`` KAT (max_time 6.062).
* ``gss_oracle``      - PINNED against ``golden_section_search.gss`` evaluation sequences.
* ``vad_oracle``      - PARITY UNPINNED for the detector arithmetic: the reference's VADs
  live in third-party wheels (webrtcvad, auditok==0.1.5, silero) that are absent here
  and no reference test runs a real VAD.  The energy/zero-crossing rule is this
  repo's own definition; only its shape/label/chunk contract is pinned
  to ``speech_transformers.py:155-183``.
"""
