"""Informational only (not a bench arm): the SAME torch ops the reference dispatches (oracle restatement = stock ATen / cuDNN
eager kernels) run on the B200 itself, fwd + loss + bwd of the C2 batch, fp32 without TF32, TF32, and bf16 autocast.
This is the number our kernels have to beat on the GPU; the contractual reference arm of bench.py is the CPU path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import fusion_fpn_oracle as O

sd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.make_state_dict(seed=1234).items()}
batch = {k: v.cuda() for k, v in O.synthetic_batch(8, 32, 128, 128, 320, 128, seed=1234).items()}
torch.backends.cudnn.benchmark = True
keys = set(O.param_keys(sd))


def run(label, autocast=None, tf32=False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    def step():
        if autocast is None:
            return O.loss_and_grads(sd, batch)
        # autocast around the network only: binary_cross_entropy refuses to run under autocast
        work = {k: (v.clone().requires_grad_(True) if torch.is_tensor(v) and v.is_floating_point() and k in keys else v) for k, v in sd.items()}
        with torch.autocast('cuda', dtype=autocast):
            out = O.fpn_hybrid_fusion_forward(work, batch, 'relative_2d_max', 'slo', True, True, None, None)
        loss = O.mix_loss(out['prediction'].float(), batch['mask'])
        torch.autograd.grad(loss, [work[k] for k in O.param_keys(sd)])
        return (loss.detach(),)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        loss = step()[0]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f'{label}: {ms:.2f} ms per fwd+loss+bwd of batch 8 -> {8 / ms * 1e3:.1f} samples/s (loss {float(loss):.4f})', flush=True)


if not os.environ.get('ONLY_BF16'):
    run('torch eager fp32 (TF32 off)')
    run('torch eager TF32', tf32=True)
run('torch eager bf16 autocast', autocast=torch.bfloat16, tf32=True)
