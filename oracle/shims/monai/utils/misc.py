def ensure_tuple_rep(tup, dim):
    if isinstance(tup, (tuple, list)):
        if len(tup) == dim:
            return tuple(tup)
        raise ValueError("sequence length mismatch")
    return (tup,) * dim
