#!/bin/bash
# run-to-run reproducibility of the smoke step: the two printed loss triples must be identical
python __graft_entry__.py smoke 2>&1 | tail -n 1
python __graft_entry__.py smoke 2>&1 | tail -n 1
