"""Network descriptors: the host-side mirror of the reference's `module.py`.

In the reference, `generator_gatedcnn` (module.py:148-185) and `discriminator` (module.py:188-213) are
graph-building callables injected into `CycleGAN(...)` (model.py:9,14-15).  Here they are *descriptors*:
objects that name the native kernel graph libcgvc.so runs for that network.  The engine implements exactly
these two architectures; passing anything else to `CycleGAN` raises, rather than silently running a
different model.
"""
from __future__ import annotations


class _NetDescriptor:
    def __init__(self, name, kind, layers):
        self.__name__ = name
        self.kind = kind
        self.layers = layers

    def __call__(self, inputs=None, reuse=False, scope_name=None):
        raise TypeError(
            "%s is a network descriptor for the native engine, not a TensorFlow graph builder; "
            "use CycleGAN(...).test(inputs, direction) to run it" % self.__name__)

    def __repr__(self):
        return "<native %s descriptor: %s>" % (self.kind, self.__name__)


# (name, kernel, stride, filters) -- module.py:161-183
generator_gatedcnn = _NetDescriptor(
    "generator_gatedcnn", "generator",
    [("h1_conv|h1_conv_gates + GLU", 15, 1, 128),
     ("downsample1d_block1", 5, 2, 256), ("downsample1d_block2", 5, 2, 512)]
    + [("residual1d_block%d" % i, 3, 1, 1024) for i in range(1, 7)]
    + [("upsample1d_block1 (+pixel shuffle)", 5, 1, 1024), ("upsample1d_block2 (+pixel shuffle)", 5, 1, 512),
       ("o1_conv", 15, 1, 24)])

# module.py:201-211
discriminator = _NetDescriptor(
    "discriminator", "discriminator",
    [("h1_conv|h1_conv_gates + GLU", (3, 3), (1, 2), 128),
     ("downsample2d_block1", (3, 3), (2, 2), 256), ("downsample2d_block2", (3, 3), (2, 2), 512),
     ("downsample2d_block3", (6, 3), (1, 2), 1024), ("dense + sigmoid", None, None, 1)])
