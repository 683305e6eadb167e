"""Row f-2 at full size on the B200: GPTQ-calibrate the whole ViT-H image encoder (32 blocks, 128 linears)
on one synthetic image with the solver's heavy parts on the device (Hessian: samq_syrk_f32_fwd, rounding
loop: samq_gptq_block_fwd), pack it, and run the packed encoder on the fused kernels.

    python tests/runs/gptq_vith_gpu.py [out.json]          # ~minutes on one B200

Reports the time of each phase and how close the int4 encoder's embeddings are to the fp16 encoder's and to
round-to-nearest at the same bit width (GPTQ has to beat RTN on the layer outputs it was calibrated on)."""
import copy
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import sam_quantization_b200 as sq  # noqa: E402
from sam_quantization_b200 import _lib  # noqa: E402
from sam_quantization_b200 import gptq as G  # noqa: E402
from sam_quantization_b200.image_encoder import build_image_encoder  # noqa: E402


def cos(a, b):
    return torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/gptq_vith_gpu.json"
    model = os.environ.get("SAMQ_GPTQ_MODEL", "vit_h")
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    enc = build_image_encoder(model).half().to(dev).eval()
    with torch.no_grad():
        for n, p in enc.named_parameters():
            if "rel_pos" in n or n.endswith("pos_embed"):
                p.copy_(torch.randn(p.shape, device=dev) * 0.02)
    img = torch.randn(1, 3, 1024, 1024, device=dev, generator=torch.Generator(dev).manual_seed(1)).half()
    res = {"model": model, "wbits": 4, "groupsize": 128}
    with torch.no_grad():
        y_fp = enc(img).float()
        # round-to-nearest at the same format, for comparison
        rtn = copy.deepcopy(enc)
        for m in rtn.modules():
            if isinstance(m, torch.nn.Linear) and m.in_features % 128 == 0 and m.weight.shape[0] >= 256:
                w = m.weight.data.float()
                q = G.Quantizer()
                q.configure(4, perchannel=True, sym=False, mse=False)
                for c in range(0, w.shape[1], 128):
                    q.find_params(w[:, c:c + 128], weight=True)
                    w[:, c:c + 128] = q.quantize(w[:, c:c + 128])
                m.weight.data = w.half()
        y_rtn = rtn(img).float()
        del rtn
        launches = _lib.launch_count()
        torch.cuda.synchronize()
        t0 = time.time()
        qs = G.encoder_sequential(enc, [img], wbits=4, nsamples=1, groupsize=128)
        torch.cuda.synchronize()
        res["calibrate_s"] = round(time.time() - t0, 2)
        res["calibrate_libsamq_launches"] = int(_lib.launch_count() - launches)
        res["layers"] = len(qs)
        y_rounded = enc(img).float()              # the solver leaves the rounded weights in the model
        t0 = time.time()
        G.encoder_pack(enc, qs, 4, 128)
        res["pack_s"] = round(time.time() - t0, 2)
        sq.make_quant_attn(enc)
        sq.make_fused_mlp(enc)
        enc = enc.to(dev).eval()
        y_q = enc(img).float()
        torch.cuda.synchronize()
    res["cosine_packed_vs_rounded_eager"] = cos(y_q, y_rounded)
    res["maxabs_packed_vs_rounded_eager"] = float((y_q - y_rounded).abs().max())
    res["cosine_gptq_vs_fp16"] = cos(y_q, y_fp)
    res["cosine_rtn_vs_fp16"] = cos(y_rtn, y_fp)
    res["rel_err_gptq"] = float((y_q - y_fp).norm() / y_fp.norm())
    res["rel_err_rtn"] = float((y_rtn - y_fp).norm() / y_fp.norm())
    # the reference's own solver (gptq.py, torch ops on the same GPU, one column at a time) on the four
    # layer shapes of a ViT-H block with the same kind of input, for the record
    try:
        from oracle import make_ref

        make_ref.add_to_path()
        import gptq as ref_gptq            # the staged reference module (oracle/_ref/gptq.py)

        D = enc.blocks[0].norm1.weight.numel()
        shapes = {"qkv": (D, 3 * D), "proj": (D, D), "lin1": (D, 4 * D), "lin2": (4 * D, D)}
        t_ref = t_ours = 0.0
        for name, (k, n) in shapes.items():
            x = (torch.randn(1, 4096, k, device=dev, generator=torch.Generator(dev).manual_seed(k + n)) *
                 torch.linspace(0.3, 2.0, k, device=dev)).half()
            w = torch.randn(n, k, device=dev, generator=torch.Generator(dev).manual_seed(n)) * 0.02
            for which in ("ref", "ours"):
                lin = torch.nn.Linear(k, n, bias=False).to(dev)
                lin.weight.data = w.clone()
                if which == "ref":
                    s_ = ref_gptq.GPTQ(lin)
                    s_.quantizer = ref_gptq.Quantizer()
                else:
                    s_ = G.GPTQ(lin)
                    s_.quantizer = G.Quantizer()
                s_.quantizer.configure(4, perchannel=True, sym=False, mse=False)
                torch.cuda.synchronize()
                t0 = time.time()
                s_.add_batch(x, None)
                s_.fasterquant(percdamp=0.01, groupsize=128)
                torch.cuda.synchronize()
                dt = time.time() - t0
                xf = x[0].float()
                out_err = float((xf @ (lin.weight.data.float() - w).t()).norm() / (xf @ w.t()).norm())
                if which == "ref":
                    t_ref += dt
                    q_ref = lin.weight.data.clone()
                    res[f"{name}_layer_output_rel_err_reference_solver"] = out_err
                else:
                    t_ours += dt
                    flips = ((lin.weight.data - q_ref).abs() > 1e-6).float().mean().item()
                    res[f"{name}_rounded_weight_mismatch_vs_reference_solver"] = flips
                    res[f"{name}_layer_output_rel_err_device_solver"] = out_err
        res["one_block_reference_solver_s"] = round(t_ref, 3)
        res["one_block_device_solver_s"] = round(t_ours, 3)
    except Exception as ex:      # the staged reference is optional
        res["reference_solver"] = f"not run: {ex!r}"
    print(json.dumps(res, indent=1))
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
