"""Generates tests/golden/gptq_*.npz from the REFERENCE's own solver (gptq.py: GPTQ.add_batch /
fasterquant, Quantizer.find_params) imported from /root/reference and run on the CPU.  Run once
in the build container (the reference does not travel to the GPU box); the committed fixtures pin
sam_quantization_b200/gptq.py.

    python tests/golden/make_gptq_fixtures.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: rows, cols, tokens per call, calls, bits, groupsize, blocksize, actorder, sym, dead column
    "g64": dict(rows=40, cols=192, tokens=160, calls=2, bits=4, groupsize=64, blocksize=64, actorder=False, sym=False, dead=None),
    "g32_inblock": dict(rows=24, cols=192, tokens=200, calls=1, bits=4, groupsize=32, blocksize=128, actorder=False, sym=False, dead=5),
    "perrow_b3": dict(rows=32, cols=160, tokens=180, calls=3, bits=3, groupsize=-1, blocksize=64, actorder=False, sym=True, dead=None),
    "actorder_b8": dict(rows=16, cols=128, tokens=150, calls=1, bits=8, groupsize=64, blocksize=64, actorder=True, sym=False, dead=None),
}


def main():
    sys.path.insert(0, REF)
    torch.cuda.synchronize = lambda *a, **k: None        # gptq.py:157 calls it unconditionally
    import gptq as ref                                    # noqa: E402  (the reference module)

    torch.set_num_threads(1)
    for name, c in CASES.items():
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        lin = nn.Linear(c["cols"], c["rows"], bias=False)
        lin.weight.data = torch.randn(c["rows"], c["cols"], generator=g) * 0.1
        W0 = lin.weight.data.clone()
        xs = []
        for _ in range(c["calls"]):
            x = torch.randn(1, c["tokens"], c["cols"], generator=g) * torch.linspace(0.2, 2.0, c["cols"])
            if c["dead"] is not None:
                x[..., c["dead"]] = 0
            xs.append(x)
        solver = ref.GPTQ(lin)
        solver.quantizer = ref.Quantizer()
        solver.quantizer.configure(c["bits"], perchannel=True, sym=c["sym"], mse=False)
        for x in xs:
            solver.add_batch(x, None)
        H = solver.H.clone()
        scale, zero = solver.fasterquant(blocksize=c["blocksize"], percdamp=0.01, groupsize=c["groupsize"],
                                         actorder=c["actorder"])
        np.savez_compressed(os.path.join(HERE, f"gptq_{name}.npz"), W0=W0.numpy(),
                            X=torch.cat(xs).numpy().astype(np.float32), H=H.numpy(), Q=lin.weight.data.numpy(),
                            scale=scale.numpy(), zero=zero.numpy(),
                            cfg=np.array([c["bits"], c["groupsize"], c["blocksize"], int(c["actorder"]), int(c["sym"])]))
        print(name, "Q", tuple(lin.weight.shape), "scale", tuple(scale.shape))

    # Quantizer alone: min/max, symmetric, and the mse grid search
    g = torch.Generator().manual_seed(77)
    w = torch.randn(48, 96, generator=g) * torch.linspace(0.05, 1.0, 48).unsqueeze(1)
    w[3] = 0
    out = {"w": w.numpy()}
    for tag, kw in {"asym": dict(sym=False, mse=False), "sym": dict(sym=True, mse=False),
                    "mse": dict(sym=False, mse=True), "tensor": dict(sym=False, mse=False, perchannel=False)}.items():
        q = ref.Quantizer()
        q.configure(4, perchannel=kw.pop("perchannel", True), **kw)
        q.find_params(w.clone(), weight=True)
        out[f"{tag}_scale"], out[f"{tag}_zero"] = q.scale.numpy(), q.zero.numpy()
        out[f"{tag}_fq"] = q.quantize(w).numpy()
    np.savez_compressed(os.path.join(HERE, "gptq_quantizer.npz"), **out)
    print("quantizer fixtures written")


if __name__ == "__main__":
    main()
