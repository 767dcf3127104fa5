"""Row-sharded forward on 2 GPUs (one process per GPU, NCCL all-reduce of the residual contribution twice per block)
against the unsharded CPU oracle.  Partial sums re-associate the f32 accumulation -> tolerance 1e-3, not bit-equality."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir, name, expect_error=None):
    for p in (os.path.join(ROOT, "lm.rs_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import lmrs_b200
    from lmrs_b200 import lmrs_file as lf
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(lmrs_b200.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    buf = lf.write_synthetic(lf.model_args(name, 1))
    if expect_error:   # shapes whose shards would split a 128-element quantization group are refused at load, on every rank
        with pytest.raises(lmrs_b200.LmrsError, match=expect_error):
            lmrs_b200.Transformer.new_sharded(buf, rank, rank, world, bytes(idt.cpu().numpy().tobytes()))
        dist.barrier()
        dist.destroy_process_group()
        return
    m, _ = lmrs_b200.Transformer.new_sharded(buf, rank, rank, world, bytes(idt.cpu().numpy().tobytes()))
    toks = np.random.default_rng(1).integers(0, m.args.vocab_size, 10)
    worst = 0.0
    if rank == 0:
        import lmrs_ref
        cpu = lmrs_ref.RefTransformer(buf)
    emb = m.get_embeddings(toks[:4])
    assert m.fill_kv_cache(emb, 0) == 4
    if rank == 0:
        ec = cpu.get_embeddings(toks[:4]); cpu.fill_kv_cache(ec, 0)
        worst = max(worst, float(np.abs(emb - ec).max()))
    for i, t in enumerate(toks[4:]):
        lg = m.forward(int(t), 4 + i)
        if rank == 0:
            worst = max(worst, float(np.abs(lg - cpu.forward(int(t), 4 + i)).max()))
    if rank == 0:
        open(os.path.join(out_dir, "worst.txt"), "w").write(repr(worst))
    dist.barrier()
    dist.destroy_process_group()


# tiny models only: on deeper random-weight models any f32 re-association flips activation-quantization codes and the
# deviation is amplified far beyond 1e-3, which is exactly what the partial sums of N > 1 do
def test_two_gpu_sharded_forward_matches_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), "tiny-llama"), nprocs=2, join=True)
    assert float(open(tmp_path / "worst.txt").read()) <= 1e-3


def test_two_gpu_rejects_shapes_that_split_a_quantization_group(tmp_path):
    """tiny-phi: att_dim 384 / 2 ranks = 192 is not a multiple of the 128-element group of Wo's input -> loud error."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), "tiny-phi", "multiple of 128"), nprocs=2, join=True)
