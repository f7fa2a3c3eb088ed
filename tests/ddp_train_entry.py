"""Helper launched by torchrun from test_gpu_entrypoints.py (2-rank NCCL training)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    import torch

    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import train
    root, log_dir = sys.argv[1], sys.argv[2]
    args = train.parse_args([], arch="GCN_FC(8,8,hids=64,depth=3)", epochs=4, dataset_processed_prefix=root, log_dir=log_dir,
                            num_workers=0, log_every=1)
    model, _ = train.run_exp(args)
    # replicas must be bit-identical after training (same averaged gradients, same optimiser state)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(torch.distributed.get_world_size())]
    torch.distributed.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], g) for g in gathered), "replicas diverged"
    torch.distributed.destroy_process_group()
