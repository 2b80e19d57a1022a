"""Drop-in import paths of the reference (README.md:29-33 of IoBT-VISTEC/OCTAve), served by octave_b200."""
