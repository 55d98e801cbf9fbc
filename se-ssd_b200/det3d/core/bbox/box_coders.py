"""Box coders (reference: det3d/core/bbox/box_coders.py:32-106)."""
from . import box_np_ops, box_torch_ops


class GroundBox3dCoder(object):
    def __init__(self, linear_dim=False, vec_encode=False, n_dim=7, norm_velo=False):
        self.linear_dim, self.vec_encode, self.norm_velo, self.n_dim = linear_dim, vec_encode, norm_velo, n_dim

    @property
    def code_size(self):
        return self.n_dim + 1 if self.vec_encode else self.n_dim

    def encode(self, boxes, anchors):
        return box_np_ops.second_box_encode(boxes, anchors, self.vec_encode, self.linear_dim)

    def decode(self, rel_codes, anchors):
        return box_np_ops.second_box_decode(rel_codes, anchors, self.vec_encode, self.linear_dim)


class GroundBox3dCoderTorch(GroundBox3dCoder):
    def encode_torch(self, boxes, anchors):
        return box_torch_ops.second_box_encode(boxes, anchors, self.vec_encode, self.linear_dim)

    def decode_torch(self, boxes, anchors):
        return box_torch_ops.second_box_decode(boxes, anchors, self.vec_encode, self.linear_dim)
