"""Synthetic loader batches of the shape the reference trainer consumes; used by bench.py, the
tests and the oracle (there is no network for the real dataset)."""
import torch


def synthetic_batch(B: int, L: int, S: int, seed: int = 1234):
    """Synthetic loader batch as SURVEY.md 8(d): frames/cad in [-1,1], raw action rows with a valid
    command/parameter structure, first row all-zero.  L = T + 1 loaded steps."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.randn(B, L, 1, S, S, generator=g).clamp_(-1, 1)
    cad = torch.randn(B, 1, S, S, generator=g).clamp_(-1, 1)
    actions = -torch.ones(B, L, 7)
    cmd = torch.randint(0, 5, (B, L), generator=g)
    actions[..., 0] = cmd.float()
    xy = torch.randint(0, 1000, (B, L, 2), generator=g).float()
    key = (torch.randint(0, 20, (B, L), generator=g) * 50).float()
    times = torch.tensor([-1.0, 0.0, 200.0, 400.0])[torch.randint(0, 4, (B, L), generator=g)]
    scroll = (torch.randint(0, 2, (B, L), generator=g) * 500).float()
    typed = torch.randint(0, 1000, (B, L), generator=g).float()
    m0, m1, m2, m3 = (cmd == 0), (cmd == 1), (cmd == 2), (cmd == 3)
    actions[..., 1] = torch.where(m0, xy[..., 0], actions[..., 1])
    actions[..., 2] = torch.where(m0, xy[..., 1], actions[..., 2])
    actions[..., 3] = torch.where(m1, key, actions[..., 3])
    actions[..., 4] = torch.where(m1, times, actions[..., 4])
    actions[..., 5] = torch.where(m2, scroll, actions[..., 5])
    actions[..., 6] = torch.where(m3, typed, actions[..., 6])
    actions[:, 0, :] = 0.0
    return {"frames": frames, "actions": actions, "cad_image": cad}
