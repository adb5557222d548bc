"""Seeded adversarial FASTA/FASTQ generators for the parity tests."""
import numpy as np


def random_fasta(seed, n_records=50, max_len=3000, widths=(60, 70, 80), crlf_prob=0.2, odd_line_prob=0.15,
                 blank_prob=0.05, long_line_prob=0.05, no_trailing_newline=False, header_tab_prob=0.2,
                 lowercase_prob=0.2, lead_garbage=False):
    rng = np.random.default_rng(seed)
    out = []
    if lead_garbage:
        out.append(b"\n\n")
    crlf_file = rng.random() < crlf_prob
    eol = b"\r\n" if crlf_file else b"\n"
    alphabet = np.frombuffer(b"ACGTNacgtnRYKMSWBDHV", dtype=np.uint8)
    for i in range(n_records):
        name = b"seq%d_%d" % (seed, i)
        r = rng.random()
        if r < header_tab_prob:
            hdr = b">" + name + b"\tdesc with tab"
        elif r < 0.6:
            hdr = b">" + name + b" some description %d" % i
        else:
            hdr = b">" + name
        out.append(hdr + eol)
        L = int(rng.integers(0, max_len + 1))
        if rng.random() < lowercase_prob:
            seq = alphabet[rng.integers(0, alphabet.size, size=L)]
        else:
            seq = alphabet[rng.integers(0, 4, size=L)]
        w = int(rng.choice(widths))
        if rng.random() < long_line_prob:
            w = max(1, L)            # unwrapped record
        lines = [seq[k:k + w].tobytes() for k in range(0, L, w)]
        if lines and rng.random() < odd_line_prob:
            j = int(rng.integers(0, len(lines)))
            cut = int(rng.integers(1, max(2, len(lines[j]))))
            extra = lines[j][cut:]
            lines[j] = lines[j][:cut]
            if extra:
                lines.insert(j + 1, extra)
        for ln in lines:
            out.append(ln + eol)
        if rng.random() < blank_prob:
            out.append(eol)
    data = b"".join(out)
    if no_trailing_newline and data.endswith(eol):
        data = data[:-len(eol)]
    return data


def random_fastq(seed, n_reads=200, max_len=300, crlf=False, no_trailing_newline=False, partial_tail=0):
    rng = np.random.default_rng(seed)
    eol = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(n_reads):
        L = int(rng.integers(1, max_len + 1))
        seq = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, size=L)].tobytes()
        q = rng.integers(33, 75, size=L).astype(np.uint8).tobytes()   # may contain '@' and '+'
        r = rng.random()
        if r < 0.3:
            name = b"@read%d" % i
        elif r < 0.7:
            name = b"@read%d 1:N:0:ACGT extra" % i
        else:
            name = b"@read%d\twith tab" % i
        plus = b"+" if rng.random() < 0.7 else b"+" + name[1:]
        out += [name + eol, seq + eol, plus + eol, q + eol]
    lines = out
    if partial_tail:
        lines = lines + [b"@tail x" + eol, b"ACGT" + eol, b"+" + eol][:partial_tail]
    data = b"".join(lines)
    if no_trailing_newline and data.endswith(eol):
        data = data[:-len(eol)]
    return data


def random_queries(rows, nq, seed, max_len=None):
    """(row_id, s, e) uniform over records with slen > 0; includes whole-record queries."""
    rng = np.random.default_rng(seed)
    slen = np.asarray(rows["slen"], dtype=np.int64)
    cand = np.nonzero(slen > 0)[0]
    if cand.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)
    rid = cand[rng.integers(0, cand.size, size=nq)]
    s = (rng.random(nq) * slen[rid]).astype(np.int64)
    e = s + 1 + (rng.random(nq) * (slen[rid] - s)).astype(np.int64)
    e = np.minimum(e, slen[rid])
    if max_len:
        e = np.minimum(e, s + max_len)
    whole = rng.random(nq) < 0.1
    s[whole] = 0
    e[whole] = slen[rid][whole]
    return rid.astype(np.int64), s, e
