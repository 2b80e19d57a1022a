from octave_b200.network import AdversarialAttentionGate, GlobalAveragePooling2D  # noqa: F401
