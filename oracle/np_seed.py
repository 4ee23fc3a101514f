"""Vectorised restatement of numpy's SeedSequence -> PCG64 seeding (TEST INFRASTRUCTURE; pinned against numpy itself in
tests/test_oracle_seeding.py).

numpy is a third-party dependency of the reference (Randomizer.reset, components/tetromino_randomizer.py:40-43 calls
`np.random.default_rng(seed)` = `Generator(PCG64(SeedSequence(seed)))`; poetry.lock pins numpy 2.2.4).  Published algorithm
(numpy/random/bit_generator.pyx, SeedSequence): the seed's 32-bit words are hashed into a pool of 4 words (hashmix / mix), 8
output words are generated from the pool (`generate_state(4, uint64)`), and PCG64 is seeded with
`pcg_setseq_128_srandom_r(state = w0:w1, seq = w2:w3)` (numpy/random/src/pcg64/pcg64.h).
"""
import numpy as np

INIT_A, MULT_A = 0x43B0D7E5, 0x931E8875
INIT_B, MULT_B = 0x8B51F9DD, 0x58F38DED
MIX_MULT_L, MIX_MULT_R = 0xCA01F9DD, 0x4973F715
XSHIFT = 16
M32 = 0xFFFFFFFF
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645


def seed_words(seeds):
    """uint64 seeds [n] -> generate_state(4, uint64) of SeedSequence(seed) as uint64[n, 4] (vectorised over seeds)."""
    seeds = np.asarray(seeds, dtype=np.uint64)
    u32 = np.uint32
    ent = [(seeds & np.uint64(M32)).astype(u32), (seeds >> np.uint64(32)).astype(u32)]   # little-endian words; hi == 0 hashes like "absent"
    zero = np.zeros_like(ent[0])
    hc = INIT_A                                   # hash_const is a scalar sequence, the same for every seed

    def hashmix(v):
        nonlocal hc
        v = v ^ u32(hc)
        hc = (hc * MULT_A) & M32
        v = v * u32(hc)
        return v ^ (v >> u32(XSHIFT))

    def mix(x, y):
        r = u32(MIX_MULT_L) * x - u32(MIX_MULT_R) * y
        return r ^ (r >> u32(XSHIFT))

    with np.errstate(over="ignore"):
        pool = [hashmix(ent[i] if i < 2 else zero) for i in range(4)]
        for i_src in range(4):
            for i_dst in range(4):
                if i_src != i_dst:
                    pool[i_dst] = mix(pool[i_dst], hashmix(pool[i_src]))
        hb = INIT_B
        words = []
        for i in range(8):
            d = pool[i % 4] ^ u32(hb)
            hb = (hb * MULT_B) & M32
            d = d * u32(hb)
            words.append(d ^ (d >> u32(XSHIFT)))
    out = np.empty((len(seeds), 4), np.uint64)
    for k in range(4):
        out[:, k] = words[2 * k].astype(np.uint64) | (words[2 * k + 1].astype(np.uint64) << np.uint64(32))
    return out


def pcg64_from_words(w):
    """pcg_setseq_128_srandom_r on one row of seed_words -> (state_hi, state_lo, inc_hi, inc_lo); scalar (python ints)."""
    m128 = (1 << 128) - 1
    initstate = (int(w[0]) << 64) | int(w[1])
    initseq = (int(w[2]) << 64) | int(w[3])
    inc = ((initseq << 1) | 1) & m128
    state = (0 * PCG_MULT + inc) & m128
    state = (state + initstate) & m128
    state = (state * PCG_MULT + inc) & m128
    m64 = (1 << 64) - 1
    return state >> 64, state & m64, inc >> 64, inc & m64
