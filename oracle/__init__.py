"""CPU oracle for the CT-GAN training step — TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a CPU restatement (PyTorch-CPU, fp64 master / fp32) of the
reference's critic/generator training step:

  TG = /root/reference/CT-GANs/tensorflow_generative_model
  TG/tflib/__init__.py, TG/tflib/ops/{conv2d,deconv2d,linear,batchnorm,cond_batchnorm}.py
  TG/CT_gan_mnist.py:39-177, TG/CT_gan_cifar.py:47-154, TG/CT_gan_cifar_resnet.py:67-338

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it, and only as the checker / the timed CPU
baseline.  Nothing under `ctgan_b200/` imports it.

Pinning status
--------------
The reference ships NO tests, golden vectors or fixtures, and
its arithmetic lives in the third-party dependency `tensorflow==1.2.1`
(README.md:3), which is absent from /root/reference and not installable here.
What CAN be pinned is pinned: `oracle/ref_harness.py` executes the reference's
OWN tflib / script source (translated py2→py3 into `oracle/_ref/`, never
committed) against `oracle/tf_shim` — a restatement of the documented TF-1.x
semantics of the ~30 `tf.*` calls on the path — and `tests/test_oracle_vs_reference.py`
checks this restatement against those runs; `tests/golden/*.npz` were generated
from it (script: `tests/golden/make_golden.py`).  TensorFlow's own kernels were
never executed, so for the TF-internal arithmetic (SAME padding, conv2d_transpose
cropping, fused_batch_norm, Adam epsilon placement) parity is **unpinned**:
it follows TF's published behaviour as listed in SURVEY.md §8(c).
"""
