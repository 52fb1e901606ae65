"""One process, several GPUs (BASELINE.json north_star: "multi-field and multi-level calls are partitioned across the 8 GPUs of
one box by level, each GPU holding a replicated gset"): c_ezb200_set_devices(n) fans a host-array batch call out over n devices.
The result must equal the single-device call bit for bit (every device builds the same gridset and runs the same kernels)."""
import numpy as np
import pytest

from librmn_b200 import configs

pytestmark = pytest.mark.gpu


def test_fan_out_equals_single_device(ezb, cuda_ok):
    import torch
    ndev = torch.cuda.device_count()
    for mk, nf in ((lambda: configs.c5(ezb, 0.6, 0.4, "cubic"), 13), (lambda: configs.c2(ezb, 0.6, 0.4, "cubic"), 9), (lambda: configs.c3(ezb, 4, "cubic"), 8)):
        c = mk()
        c.defset()
        if c.vector:
            uv = [c.source_uv(k) for k in range(nf)]
            a = [np.stack([x for x, _ in uv]), np.stack([y for _, y in uv])]
        else:
            a = [np.stack([c.source(k) for k in range(nf)])]
        res = {}
        for n in (1, min(2, ndev), ndev):
            used = ezb.set_devices(n)
            assert used == min(n, ndev)
            try:
                outs = [np.zeros((nf,) + c.dst_shape, np.float32) for _ in a]
                rc = ezb.ezuvint_batch(outs[0], outs[1], a[0], a[1], nf) if c.vector else ezb.ezsint_batch(outs[0], a[0], nf)
                assert rc >= 0
                res[n] = (rc, outs)
            finally:
                ezb.set_devices(1)
        for n, (rc, outs) in res.items():
            assert rc == res[1][0], (c.name, n)
            for o, o1 in zip(outs, res[1][1]):
                assert np.array_equal(o, o1), "%s: %d devices differ from one" % (c.name, n)
    print("devices visible:", ndev)
