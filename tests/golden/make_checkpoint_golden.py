"""Generate tests/golden/ref_checkpoint.pth with the REFERENCE's own save_checkpoint (det3d/torchie/trainer/checkpoint.py:186-214) and
check, here where the reference exists, that (a) the reference's load_checkpoint reads a file written by OUR save_checkpoint and
restores identical tensors, (b) our load_checkpoint reads the reference-written file.  The committed .pth is what
tests/test_host_logic.py::test_checkpoint_reads_reference_written_file loads on boxes without /root/reference.

    python tests/golden/make_checkpoint_golden.py
The reference file is imported where it lies; its unavailable imports (torchvision, terminaltables, det3d.torchie) are stubbed.
"""
import importlib.util
import os
import sys
import tempfile
import types
from collections import OrderedDict

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
from cases import checkpoint_model  # noqa: E402


def load_reference_checkpoint_module():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    det3d = stub("det3d")
    det3d.torchie = stub("det3d.torchie", mkdir_or_exist=lambda d, mode=0o777: os.makedirs(d, exist_ok=True) if d else None)
    stub("terminaltables", AsciiTable=object)
    if "torchvision" not in sys.modules:
        try:
            import torchvision  # noqa: F401
        except Exception:
            stub("torchvision")
    pkg = stub("refckpt")
    pkg.__path__ = []
    stub("refckpt.utils", get_dist_info=lambda: (0, 1))
    spec = importlib.util.spec_from_file_location("refckpt.checkpoint", os.path.join(REF, "det3d/torchie/trainer/checkpoint.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["refckpt.checkpoint"] = m
    spec.loader.exec_module(m)
    return m


def main():
    ref = load_reference_checkpoint_module()
    model = checkpoint_model(seed=7)

    class Wrapped(nn.Module):            # DataParallel-style wrapper: the reference unwraps `.module` (:205-206)
        def __init__(self, m):
            super().__init__()
            self.module = m

    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    out = os.path.join(HERE, "ref_checkpoint.pth")
    ref.save_checkpoint(Wrapped(model), out, optimizer=opt, meta={"epoch": 3, "iter": 1234})
    # a second fixture in the DataParallel key style ("module." prefixes, :162-163) as a bare OrderedDict (:155-156)
    out2 = os.path.join(HERE, "ref_checkpoint_module_prefix.pth")
    torch.save(OrderedDict(("module." + k, v) for k, v in ref.weights_to_cpu(model.state_dict()).items()), out2)

    sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
    spec = importlib.util.spec_from_file_location("ours_ckpt", os.path.join(ROOT, "se-ssd_b200/det3d/torchie/trainer/checkpoint.py"))
    ours = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ours)
    # (a) reference loader reads OUR file
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "ours.pth")
        ours.save_checkpoint(model, f, optimizer=opt, meta={"epoch": 3, "iter": 1234})
        fresh = checkpoint_model(seed=8)
        ck = ref.load_checkpoint(fresh, f, map_location="cpu", strict=True)
        assert ck["meta"] == {"epoch": 3, "iter": 1234} and "optimizer" in ck
        for (k, a), (_k, b) in zip(model.state_dict().items(), fresh.state_dict().items()):
            assert torch.equal(a, b), k
    # (b) our loader reads the reference's files
    for path in (out, out2):
        fresh = checkpoint_model(seed=9)
        ours.load_checkpoint(fresh, path, map_location="cpu", strict=True)
        for (k, a), (_k, b) in zip(model.state_dict().items(), fresh.state_dict().items()):
            assert torch.equal(a, b), k
    print("wrote", out, out2, "- cross-loading verified both ways")


if __name__ == "__main__":
    main()
