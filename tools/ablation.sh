#!/bin/bash
# Timing ablations of the decomposed kernels (WRONG RESULTS by construction -- never shipped): which resource does the Q1 loop wait for?
#   abl_mufu  every Box-Muller MUFU replaced by one FMUL (same dispatch slots, no XU work)
#   abl_rng   the six-instruction xorshift step replaced by one add (XU work, conversions and Weyl add stay)
#   abl_both  both
# Builds the three libraries into <pkg>/lib/variants/ (here, no GPU needed); tools/ab_variants.sh then times them on the GPU box.
python tools/build_variants.py abl_mufu=-DHW1F_ABLATE_MUFU=1 abl_rng=-DHW1F_ABLATE_RNG=1 abl_both=-DHW1F_ABLATE_RNG=1,-DHW1F_ABLATE_MUFU=1
