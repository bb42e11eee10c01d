"""The counter-based reset stream the CUDA auto-reset is specified on, on the CPU
(TEST INFRASTRUCTURE ONLY).

Stream "D2D-RNG v1": Philox4x32-10, key = (seed & 0xffffffff, seed >> 32),
counter = (global env id, episode, block, 'RSET' = 0x52534554).  Block b yields the
draws 2b and 2b+1 as 53-bit uniforms ((w0 >> 5) * 2^26 + (w1 >> 6)) / 2^53 from
words (0,1) and (2,3).  Derived draws: uniform(a,b) = a + (b-a)*u; randint(a,b) =
a + floor(u*(b-a+1)); bernoulli(p) = u < p; normal = Box-Muller cosine branch on
two consecutive draws, sqrt(-2 ln(1-u1)) cos(2 pi u2).

`PhiloxDraws` plugs this stream into oracle.ref_port.reset_tables, i.e. into the
restatement of the reference's reset (env:199-372), in the reference's draw order.
"""
import math

from . import binding as B

STREAM_RESET = 0x52534554


class PhiloxDraws:
    def __init__(self, seed, env_id, episode):
        self.key = [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]
        self.env, self.episode = env_id & 0xFFFFFFFF, episode & 0xFFFFFFFF
        self.blk = 0
        self.spare = None

    def u01(self):
        if self.spare is not None:
            v, self.spare = self.spare, None
            return v
        o = B.philox4x32_10([self.env, self.episode, self.blk, STREAM_RESET], self.key)
        self.blk += 1
        conv = lambda hi, lo: ((hi >> 5) * 67108864.0 + (lo >> 6)) * (1.0 / 9007199254740992.0)  # noqa: E731
        self.spare = conv(o[2], o[3])
        return conv(o[0], o[1])

    def uniform(self, a, b):
        return a + (b - a) * self.u01()

    py_uniform = uniform

    def randint(self, a, b):
        return min(b, a + int(math.floor(self.u01() * (b - a + 1))))

    def normal(self, mu, sd):
        u1, u2 = self.u01(), self.u01()
        return mu + sd * (math.sqrt(-2.0 * math.log(1.0 - u1)) * math.cos(2 * math.pi * u2))

    def binomial1(self, p):
        return 1 if self.u01() < p else 0
