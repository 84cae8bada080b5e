"""Run under torchrun on N GPUs: the sharded Monte-Carlo sweeps (SC fused, GRU fused, generic decoder driver) reduced with
one NCCL all-reduce must equal the single-rank run bit for bit (Philox counters = global frame index).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from neural_polar_decoder_b200 import PolarCode, construct, synth
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder, get_code
    from neural_polar_decoder_b200.sweep import mc_decoder_sweep, mc_gru_sweep, mc_sc_sweep
    ok = True
    for N, K, frames in ((1024, 512, 300001), (4096, 2048, 40003), (256, 128, 500000)):
        code = PolarCode(int(np.log2(N)), K, None, F=construct.pw_frozen_set(N, K))
        sharded = mc_sc_sweep(code, [1.5, 2.5], frames, chunk=1 << 15, seed=9)[3]          # all ranks + all-reduce
        single = mc_sc_sweep(code, [1.5, 2.5], frames, chunk=1 << 16, seed=9, rank=0, world=1)[3] if rank == 0 else None
        if rank == 0:
            same = torch.equal(sharded, single)
            ok &= same
            print("SC N=%d frames=%d world=%d: %s %s" % (N, frames, world, "equal" if same else "DIFFERENT", sharded.tolist()))
    code = get_code("Polar", "polar", 64, 22)
    net = RNN_Model('GRU', 66, 512, 1, 2, 64, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.gru_state_dict(11, 64, 512, 2, head_gain=8.0).items()})
    dec = RNN_decoder('y_input', 64, code.info_positions, onehot=True)
    sharded = mc_gru_sweep(code, net, dec, [0.0, 2.0], 50001, seed=4)[3]
    generic = mc_decoder_sweep(code, lambda y: dec.decode(net, False, y), [0.0, 2.0], 50001, chunk=7000, seed=4)[3]
    if rank == 0:
        single = mc_gru_sweep(code, net, dec, [0.0, 2.0], 50001, seed=4, rank=0, world=1)[3]
        same = torch.equal(sharded, single) and torch.equal(generic, single)
        ok &= same
        print("GRU frames=50001 world=%d: %s %s" % (world, "equal" if same else "DIFFERENT", sharded.tolist()))
        print("MULTI-GPU CHECK", "PASSED" if ok else "FAILED")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
