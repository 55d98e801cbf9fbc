"""Layer tables of the car model (shared by the runners, the weight generators and the tests)."""

# (kind, cout, ksize, stride, padding, indice_key)   det3d/models/backbones/scn.py:106-149
SPMIDDLE_LAYERS = [
    ("subm", 16, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm0"),
    ("subm", 16, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm0"),
    ("spconv", 32, (3, 3, 3), (2, 2, 2), (1, 1, 1), None),
    ("subm", 32, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm1"),
    ("subm", 32, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm1"),
    ("spconv", 64, (3, 3, 3), (2, 2, 2), (1, 1, 1), None),
    ("subm", 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm2"),
    ("subm", 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm2"),
    ("subm", 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm2"),
    ("spconv", 64, (3, 3, 3), (2, 2, 2), (0, 1, 1), None),
    ("subm", 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm3"),
    ("subm", 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm3"),
    ("subm", 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm3"),
    ("spconv", 64, (3, 1, 1), (2, 1, 1), (0, 0, 0), None),
]

# name, kind, cin, cout, k     det3d/models/necks/rpn_v1.py:135-210
SSFA_CONVS = [
    ("bottom_up_block_0.1", "conv", 128, 128, 3), ("bottom_up_block_0.4", "conv", 128, 128, 3),
    ("bottom_up_block_0.7", "conv", 128, 128, 3), ("bottom_up_block_1.0", "conv", 128, 256, 3),
    ("bottom_up_block_1.3", "conv", 256, 256, 3), ("bottom_up_block_1.6", "conv", 256, 256, 3),
    ("trans_0.0", "conv", 128, 128, 1), ("trans_1.0", "conv", 256, 256, 1),
    ("deconv_block_0.0", "deconv", 256, 128, 3), ("deconv_block_1.0", "deconv", 256, 128, 3),
    ("conv_0.0", "conv", 128, 128, 3), ("w_0.0", "conv", 128, 1, 1),
    ("conv_1.0", "conv", 128, 128, 3), ("w_1.0", "conv", 128, 1, 1),
]
