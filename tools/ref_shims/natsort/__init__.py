import re
def natsorted(seq, key=None):
    def k(s):
        s = str(s if key is None else key(s))
        return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]
    return sorted(seq, key=k)
