def __getattr__(n):
    raise RuntimeError("matplotlib stub")
