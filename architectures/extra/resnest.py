from octave_b200.network import Bottleneck, ResNestDecoder, ResNet, SplAtConv2d, Upsampling, resnest50  # noqa: F401
