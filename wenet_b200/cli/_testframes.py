"""`fsk_demod -f` (testframe mode): the sliding-window bit-error counter of reference src/fsk_demod.c:226-243, :304-343.

Host-side bookkeeping on the demodulator's output bits, nothing for the GPU: the known 100-bit frame is glibc's
rand() & 1 after srand(158324) (the same libc call here), every incoming bit shifts a 100-bit window, and a window
with fewer than 0.1 * 100 mismatches counts as one detected testframe."""
import ctypes

import numpy as np

TEST_FRAME_SIZE = 100


def tx_frame():
    libc = ctypes.CDLL(None)
    libc.srand(158324)
    return np.array([libc.rand() & 1 for _ in range(TEST_FRAME_SIZE)], dtype=np.uint8)


class TestFrames:
    __test__ = False          # not a pytest class

    def __init__(self):
        self.tx = tx_frame()
        self.window = np.zeros(TEST_FRAME_SIZE, dtype=np.uint8)
        self.frames = self.bits = self.errs = 0

    def feed(self, bits):
        """Shift `bits` (0/1 bytes, oldest first) through the window.  Returns one (bit index, errs, frames, bits,
        errs_total) tuple per detection, counters as they stand right after it (what the reference prints there)."""
        bits = np.asarray(bits, dtype=np.uint8)
        if bits.size == 0:
            return []
        cat = np.concatenate([self.window[1:], bits])
        win = np.lib.stride_tricks.sliding_window_view(cat, TEST_FRAME_SIZE)      # win[j] = window after bits[j]
        errs = (win != self.tx).sum(axis=1)
        out = []
        for j in np.nonzero(errs < 0.1 * TEST_FRAME_SIZE)[0]:
            self.frames += 1
            self.bits += TEST_FRAME_SIZE
            self.errs += int(errs[j])
            out.append((int(j), int(errs[j]), self.frames, self.bits, self.errs))
        self.window = cat[-TEST_FRAME_SIZE:].copy()
        return out

    @staticmethod
    def line(hit):
        """the stderr line of src/fsk_demod.c:337-338 (printed when -t is not given)"""
        _, errs, _, nbits, nerrs = hit
        ber = float(np.float32(nerrs) / np.float32(nbits))
        return "errs: %d FSK BER %f, bits tested %d, bit errors %d\n" % (errs, ber, nbits, nerrs)
