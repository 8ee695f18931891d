def windowed(seq, n, fillvalue=None, step=1):
    seq = list(seq)
    if n > len(seq):
        yield tuple(seq + [fillvalue] * (n - len(seq)))
        return
    i = 0
    while True:
        w = seq[i:i + n]
        if len(w) < n:
            if i - step + n < len(seq):  # leftover items not yet covered
                yield tuple(w + [fillvalue] * (n - len(w)))
            return
        yield tuple(w)
        if i + n >= len(seq):
            return
        i += step
