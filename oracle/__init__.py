"""CPU oracle for the CRW hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``radar_sounder_crw_b200``) never does: it fails loudly when its CUDA library
is missing instead of falling back to anything in here.

Contents
--------
walk_oracle.py       numpy restatement of the training walk of the reference
                     (``src/model.py:22-46``), forward in the reference's own
                     (T-2)^2 loop order and in the O(T) chain form, plus the
                     analytic backward.
labelprop_oracle.py  numpy restatement of label propagation
                     (``src/utils.py:134-161``, ``src/imported/labelprop.py:67-116``,
                     ``src/imported/maskedatt.py:151-175,232-245``).
crw_oracle.c         plain-C restatement of label propagation with a *pinned*
                     fp32 operation order (one fmaf chain per float4 of channels + xor butterfly, a
                     fixed polynomial exp) so the CUDA fp32 path can be compared
                     bit for bit; also the multi-threaded CPU baseline.
walk_torch_port.py   torch-CPU port of the reference train step used only to
                     time the CPU baseline (autograd through the encoder).
ref_shim.py          imports the *live* reference from /root/reference (only
                     exists in the build container) to pin the oracle and to
                     generate ``tests/golden/*.npz``.

Parity pin: the reference ships no golden vectors or tests,
so the oracle is pinned against outputs of the live reference code run in the
build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
"""
