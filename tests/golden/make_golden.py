"""Generates the committed golden fixtures by RUNNING THE REFERENCE'S OWN PYTHON EXPORTER (export.py, utils/io.py,
utils/quantization.py at /root/reference) on tiny seeded safetensors models.  Only runs in the authoring
container (the GPU box has no /root/reference); the outputs are committed:

    tests/golden/<name>.lmrs          the exporter's output, byte for byte
    tests/golden/<name>.src.npz       the f32 source tensors fed to the exporter (HF names)

They pin (a) the LMRS v4 layout our loader/writer assume and (b) the Q8_0/Q4_0 weight quantizers.
usage: python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch
from safetensors.torch import save_file

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (type, quantize_type, cfg)
    "ref_llama_q8": ("LLAMA", 1, dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                                      num_key_value_heads=1, vocab_size=64, max_position_embeddings=131072,
                                      rms_norm_eps=1e-5, rope_theta=500000.0)),
    "ref_gemma_q4": ("GEMMA", 2, dict(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2,
                                      num_key_value_heads=1, head_dim=128, vocab_size=64, max_position_embeddings=8192,
                                      rms_norm_eps=1e-6, rope_theta=10000.0)),
    "ref_phi_q8": ("PHI", 1, dict(hidden_size=384, intermediate_size=256, num_hidden_layers=1, num_attention_heads=4,
                                  num_key_value_heads=4, vocab_size=64, max_position_embeddings=131072,
                                  rms_norm_eps=1e-5, rope_theta=10000.0)),
}


def tensors_for(mtype, cfg, seed):
    rng = np.random.default_rng(seed)
    d, hd, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
    hs = cfg.get("head_dim", d // cfg["num_attention_heads"])
    att, kv = cfg["num_attention_heads"] * hs, cfg["num_key_value_heads"] * hs
    def w(*shape):   # values on a 1/4096 grid so the npz compresses; still exercises every code
        return torch.from_numpy((np.round(rng.standard_normal(shape) * 0.25 * 4096) / 4096).astype(np.float32))
    t = {"model.embed_tokens.weight": w(cfg["vocab_size"], d), "model.norm.weight": 1 + w(d)}
    for l in range(L):
        p = f"model.layers.{l}."
        t[p + "input_layernorm.weight"] = 1 + w(d)
        t[p + "post_attention_layernorm.weight"] = 1 + w(d)
        if mtype == "PHI":
            t[p + "self_attn.qkv_proj.weight"] = w(att + 2 * kv, d)
            t[p + "mlp.gate_up_proj.weight"] = w(2 * hd, d)
        else:
            t[p + "self_attn.q_proj.weight"] = w(att, d)
            t[p + "self_attn.k_proj.weight"] = w(kv, d)
            t[p + "self_attn.v_proj.weight"] = w(kv, d)
            t[p + "mlp.gate_proj.weight"] = w(hd, d)
            t[p + "mlp.up_proj.weight"] = w(hd, d)
        t[p + "self_attn.o_proj.weight"] = w(d, att)
        t[p + "mlp.down_proj.weight"] = w(d, hd)
        if mtype == "GEMMA":
            t[p + "pre_feedforward_layernorm.weight"] = w(d)
            t[p + "post_feedforward_layernorm.weight"] = w(d)
    if mtype == "PHI":
        t["lm_head.weight"] = w(cfg["vocab_size"], d)
    return t


def main():
    for i, (name, (mtype, qt, cfg)) in enumerate(CASES.items()):
        with tempfile.TemporaryDirectory() as tmp:
            t = tensors_for(mtype, cfg, 100 + i)
            save_file(t, os.path.join(tmp, "model.safetensors"))
            json.dump(cfg, open(os.path.join(tmp, "config.json"), "w"))
            out = os.path.join(tmp, "out")
            subprocess.check_call([sys.executable, os.path.join(REF, "export.py"), "--files", os.path.join(tmp, "model.safetensors"),
                                   "--config", os.path.join(tmp, "config.json"), "--save-path", out, "--type", mtype,
                                   "--quantize", "--quantize-type", str(qt)], cwd=REF, stdout=subprocess.DEVNULL)
            data = open(out + ".lmrs", "rb").read()
            open(os.path.join(HERE, name + ".lmrs"), "wb").write(data)
            np.savez_compressed(os.path.join(HERE, name + ".src.npz"), **{k: v.numpy() for k, v in t.items()})
            print(name, len(data), "bytes")


if __name__ == "__main__":
    main()
