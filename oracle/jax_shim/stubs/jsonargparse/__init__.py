def CLI(*a, **k):
    raise NotImplementedError
