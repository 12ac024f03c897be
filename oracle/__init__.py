"""CPU oracle for the `jaeger predict` hot path  --  TEST INFRASTRUCTURE ONLY.

A plain NumPy / torch-fp32 restatement of the reference algorithm for every stage of the
path (windowing, encoding, conv stack, per-contig aggregation, prophage region calling).
Each function cites the reference file:line it follows (paths relative to
/root/reference/src/jaeger).

Who may import this package: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`, and there only as the checker or the timed CPU
baseline.  Nothing under `jaeger_b200/` imports it; the product path fails loudly when the
CUDA library is missing instead of falling back to this code.

Pinning status (see DESIGN.md "Oracle"):
  * windows, window metadata, encoder tokens, post-processing helpers, per-contig
    aggregation (modern and legacy, incl. the pickled legacy reliability model), TSV
    summaries, CRF / Viterbi decoding and score smoothing are pinned against golden vectors produced
    by importing the reference's own Python modules in the build container
    (tests/golden/make_goldens.py) and against the reference tests' known answers.
  * legacy (`default`) conv-stack logits: pinned on the reference's serialized TensorFlow graph, executed by the
    NumPy interpreter oracle/tfgraph.py (tests/golden/legacy_graph_outputs.npz, tests/test_legacy_graph_pin.py).
  * layer-list (v2) conv-stack logits: PARITY UNPINNED - no such SavedModel is vendored and TensorFlow/Keras cannot
    be installed here, so the forward pass is a restatement of nnlib/v2/layers.py.  Its layer functions are pinned one by one on the
    reference's own `call` bodies executed on a NumPy stand-in for TensorFlow (tests/golden/v2_layers.npz) and on the
    reference tests' mask / pooling / NMD known answers (tests/test_oracle_layer_known_answers.py); the whole representation
    learner equals the reference's DynamicModelBuilder._build_block run eagerly on that stand-in (tests/golden/v2_model.npz).
    What stays unpinned is Keras' own `__call__` mask hand-over (four rules restated in tests/golden/tf_standin.py).
  * flows around un-installable third-party calls are pinned by running the reference's own functions with the call
    stubbed by this package's restatement: scan_for_terminal_repeats / prophage_report (parasail), segment (ruptures,
    kneed), fragment_generator(dustmask=True) (pydustmasker), add_score_features / refine / aggregate_contig (polars)
    -- generators under tests/golden/make_*_goldens.py.
  * the third-party algorithms themselves stay PARITY UNPINNED: ruptures KernelCPD + kneed (restated from the published
    PELT / Kneedle algorithms), pydustmasker (published SDUST, oracle/dust.py), parasail's Smith-Waterman tie-breaking
    (oracle/termini.py), polars' summation order.
"""
