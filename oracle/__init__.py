"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path.

Nothing in ``voicecraft_b200`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and only as the checker / reported CPU baseline.

Modules
-------
patterns_oracle  numpy restatement of the delayed codebook pattern
                 (reference models/codebooks_patterns.py:117-176,302-352)
lm_oracle        torch-CPU fp32 restatement of the codec-LM decode path
                 (reference models/voicecraft.py:26-86,406-470,561-1439,
                 models/modules/{transformer,activation,embedding}.py)
encodec_oracle   torch-CPU fp32 restatement of EnCodec token->waveform decode
                 (audiocraft@c5157b5, un-vendored: "parity unpinned" by the
                 reference's own tests; pinned here against the structurally
                 identical transformers.models.encodec twin, see module header)

Pinning status: the reference ships no tests/golden vectors for this path
(SURVEY.md section 4).  lm_oracle/patterns_oracle are pinned against outputs of
the reference itself, imported and run in the build container by
``tests/golden/make_golden.py`` (fixtures committed under tests/golden/).
"""
