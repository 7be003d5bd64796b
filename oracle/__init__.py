"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the DL4DS conv super-resolution hot path.

This package is the *oracle* the CUDA path is checked against.  It is NOT part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  ``dl4ds_b200`` never imports ``oracle``.

PARITY UNPINNED: the reference (carlos-gg/dl4ds @ 232ae49) ships no tests, golden vectors or
known-answer fixtures for the network arithmetic, and its arithmetic lives in TensorFlow/Keras
(unpinned, "2.6+"; not installable here).  The only pins that exist are
  * the Keras ``model.summary()`` table recorded in ``notebooks/DL4DS_tutorial.ipynb:3335-3383``
    (per-layer parameter counts, total 487 511)  -> ``tests/test_oracle_structure.py``;
  * the reference's own numpy/cv2 batch construction, which DOES run here under import stubs
    (``oracle/ref_datapath.py``) -> committed fixtures ``tests/golden/datapath_*.npz``.
Everything else is a restatement of the published TF/Keras 2.x op definitions,
cross-checked between two independent statements: numpy-fp64 direct loops (``oracle/ops_np.py``)
and torch-CPU fp32 functional code (``oracle/torch_ref.py``).  Restatements added for SURVEY 8f row 3 carry
their own independent checks (none of them is the reference itself, so the header stays "unpinned"):
  * tf.image.ssim / ssim_multiscale -> window-by-window numpy fp64 evaluation of the SSIM definition, and the
    closed-form gradient against torch autograd (``tests/ssim_np.py``, ``tests/test_ssim_losses.py``);
  * BatchNormalization / LayerNormalization, DepthwiseConv2D, GELU -> direct numpy / scipy formulas
    (``tests/test_norm_cpu.py``, ``tests/test_convnext_cpu.py``);
  * tf.image.resize 'nearest' / 'bicubic' / 'bilinear' -> Pillow's float-image resize (and OpenCV for bilinear),
    independent implementations of the same definitions (``tests/test_resize_cpu.py``);
  * Conv2D 'same' -> scipy.signal.correlate2d; depth_to_space -> einops.rearrange; INTER_AREA coarsening -> OpenCV
    (``tests/test_oracle_independent.py``).
"""
