def trange(*a, **k): return range(*a)
def tqdm(x, *a, **k): return x
