"""The division-free state update of the device entropy coder (csrc/rans_gpu.cu: EncEntry,
enc_symbol_fast) restated with Python integers and checked against the plain division form
of SURVEY.md Appendix A.3 for every frequency class and the extreme states."""
import random

PREC = 16
L = 1 << 31


def entry(start, freq):
    if freq < 2:
        return dict(rcp=(1 << 64) - 1, shift=0, bias=start + (1 << PREC) - 1, cmpl=(1 << PREC) - freq)
    shift = 0
    while freq > (1 << shift):
        shift += 1
    rcp = ((1 << (shift + 63)) + freq - 1) // freq
    assert rcp < (1 << 64)
    return dict(rcp=rcp, shift=shift - 1, bias=start, cmpl=(1 << PREC) - freq)


def fast(x, e):
    q = ((x * e['rcp']) >> 64) >> e['shift']
    return x + e['bias'] + q * e['cmpl']


def plain(x, start, freq):
    return ((x // freq) << PREC) + (x % freq) + start


def test_reciprocal_update_equals_division_update():
    rng = random.Random(5)
    freqs = list(range(1, 70)) + [255, 256, 257, 4095, 4096, 4097, 32767, 32768, 32769, 65534, 65535]
    freqs += [rng.randrange(1, 1 << PREC) for _ in range(400)]
    for freq in freqs:
        start = rng.randrange(0, (1 << PREC) - freq + 1)
        e = entry(start, freq)
        x_max = ((L >> PREC) << 32) * freq          # the state is renormalised below this bound
        xs = [L, L + 1, x_max - 1, x_max - freq, x_max - freq - 1, (x_max // 2) | 1]
        xs += [rng.randrange(L, x_max) for _ in range(200)]
        xs += [k * freq + d for k in (L // freq + 1, x_max // freq - 1) for d in (-1, 0, 1)]
        for x in xs:
            if L <= x < x_max:
                assert fast(x, e) == plain(x, start, freq), (freq, start, x)
