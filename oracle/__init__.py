"""CPU oracle for the Video-GPT next-clip denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline -- never as the thing shipped.  The product path
(``videogpt_b200``) never imports this package and fails loudly when its CUDA
library is missing.

What it is: a plain PyTorch (CPU, any dtype) restatement of the reference's
algorithm for this path, each function citing the reference file:line it
follows (paths relative to the reference checkout):

* ``processor_oracle`` -- token layout, left padding, position ids, the dense
  ``[B,L,L]`` frame-block mask and the index dicts (``LVM/processor.py``).
* ``model_oracle``     -- ``LVM.frame_block_forward[_with_cfg]`` / ``LVM.forward``
  (``LVM/model.py``), ``Phi3Transformer.forward`` (``OmniGen/transformer.py``),
  the attention forward of ``LVM/transform/sdpa_transform.py`` and the Phi-3
  block arithmetic of the third-party dependency ``transformers==4.47.1``
  (pinned in the reference's ``env_nv.sh:16``; not vendored in the reference,
  so its published algorithm is restated here).
* ``scheduler_oracle`` -- ``LVMScheduler`` (``LVM/scheduler.py:119-208``).

Pinning: the reference has no tests and no golden vectors of its own
(SURVEY.md section 4), so the pins are outputs of the reference ITSELF run in
the build container (``oracle/refshim.py`` makes its unmodified modules
importable on CPU; ``tests/golden/make_golden.py`` generated the committed
fixtures).  ``tests/test_oracle_golden.py`` checks this restatement against
those fixtures on every run, and -- when ``/root/reference`` is present --
``tests/test_oracle_vs_reference.py`` re-runs the reference live.
"""
