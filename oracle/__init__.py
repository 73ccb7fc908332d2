"""CPU oracle for the gif-gan conv-GAN training step -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference has no tests/golden vectors and TensorFlow is
not installable here; see oracle/tf_ops.py.  Never imported by the product.
"""
