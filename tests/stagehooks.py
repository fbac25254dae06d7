"""Forward hooks that collect the per-stage activations of a fusion body (ModifiedUnet3D2D[Level5] and relatives) under
the names the oracle's ``stages`` dict uses: conv<l>, proj<l> (after the depth mean), conv<l>_2d, up<l>, final1.
The bodies call their blocks directly (the max-pool is fused behind the second block), so the hooks sit on the blocks."""


def attach(body):
    acts, hooks = {}, []

    def keep(name):
        def fn(_m, _i, o):
            acts[name] = (o[0] if isinstance(o, tuple) else o).detach()
        return fn

    for l in range(1, 6):
        for attr, name, pick in ((f'conv{l}', f'conv{l}', 1), (f'zdimRed{l}', f'proj{l}', -1), (f'conv{l}_2d', f'conv{l}_2d', 1)):
            if hasattr(body, attr):
                hooks.append(getattr(body, attr)[pick].register_forward_hook(keep(name)))
        if hasattr(body, f'up_concat{l}'):
            hooks.append(getattr(body, f'up_concat{l}').register_forward_hook(keep(f'up{l}')))
    if hasattr(body, 'final1'):
        hooks.append(body.final1.register_forward_hook(keep('final1')))
    return acts, hooks


# number of convolutions on the longest path from the inputs to a stage (bf16 tolerance = 1e-2 x depth)
def depth(name):
    l = int(''.join(ch for ch in name if ch.isdigit()) or 0)
    if name.startswith('conv'):
        return 5 * l
    if name.startswith('proj'):
        return 5 * l + (5 - l) + 1
    if name.startswith('up'):
        return 26 + 2 * (5 - l)
    return 35
