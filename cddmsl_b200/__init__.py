"""cddmsl_b200 — B200-native (sm_100a) implementation of CDDMSL's region-level vision-language hot path:
ROIAlign fwd/bwd, batched NMS, the CLIP box-predictor head (cosine logits + focal/CE loss) and the
caption-consistency alignment loss, behind the reference's own call signatures.  See DESIGN.md.

Importing the package does not load the CUDA library; the first op call does, and fails loudly if
`cddmsl_b200/lib/libcddmsl_b200.so` has not been built (`python -m cddmsl_b200.build`).
"""
__version__ = "0.1.0"
