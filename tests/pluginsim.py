"""What the plugin does to ONE track around the rate engine, restated for the tests (TEST INFRASTRUCTURE):
dsp_rate::on_chunk (foo_dsp_rate.cpp:130-206) and dsp_rate::flushwrite (:217-313) -- buffer the first block, predict
`add` frames backward from its first `prime` frames, push; keep the last block around; at the end of the track predict
`add` frames forward from the last `prime` frames, push, drain, and cut `drop` output frames from both ends.

The resampler and the predictor are parameters (objects with push / pull / drain, and a function with the signature of
lpclib.oracle_extrapolate2), so the same driver runs over the oracle, the compiled reference or the product."""
import numpy as np

LPC_ORDER = 32


def convert_track(make_resampler, extrapolate2, edge, chunks, nch):
    """chunks: list of float32 [n][nch] arrays (the audio_chunks of one track). edge = (add, drop, prime, inbuf).
    Returns the concatenated output frames the plugin would emit for the track."""
    add, drop, prime_len, inbuf = edge
    rate = make_resampler()
    out, outbuf = [], 1 << 16
    buf = np.zeros((add + inbuf + add, nch), np.float32)      # in_buffer_0; in_buffer_ starts `add` frames in
    have, dropped, primed = 0, 0, False

    def pull_once():
        y = rate.pull(outbuf)
        return np.array(y, copy=True)

    for chunk in chunks:                                      # on_chunk, :152-199
        cur, left = 0, chunk.shape[0]
        while True:
            if not primed:
                n = min(left, inbuf - have)
                buf[add + have:add + have + n] = chunk[cur:cur + n]
                have += n
                cur += n
                left -= n
                if have == inbuf:
                    extrapolate2(buf, add, prime_len, add, 0)                       # lpc_extrapolate_bkwd, :165
                    primed = True
                    rate.push(buf[:add + inbuf])
            if primed and left:
                if left < inbuf:
                    buf[add:add + inbuf - left] = buf[add + left:add + inbuf].copy()
                    buf[add + inbuf - left:add + inbuf] = chunk[cur:cur + left]
                else:
                    buf[add:add + inbuf] = chunk[cur + left - inbuf:cur + left]
                rate.push(chunk[cur:cur + left])
                cur += left
                left = 0
            y = pull_once()
            cut = min(drop - dropped, y.shape[0])
            dropped += cut
            y = y[cut:]
            if y.shape[0]:
                out.append(y)
            if not (left or y.shape[0]):
                break

    def cat():
        return np.concatenate(out, 0) if out else np.zeros((0, nch), np.float32)

    # flushwrite
    if not primed and not have > 2 * LPC_ORDER:               # :222-237, too short to predict from
        rate.push(buf[add:add + have])
        rate.drain()
        while True:
            y = pull_once()
            if y.shape[0] == 0:
                return cat()
            out.append(y)
    if not primed:                                            # :239-283, the whole track is in the buffer
        prime = min(have, prime_len)
        extrapolate2(buf, add, prime, add, 0)                                       # bkwd from the first `prime`
        extrapolate2(buf, add + have - prime, prime, 0, add)                        # fwd from the last `prime`
        rate.push(buf[:add + have + add])
        dropped = 0
        cut_front = True
    else:                                                     # :286-311, the last block is in the buffer
        extrapolate2(buf, add + inbuf - prime_len, prime_len, 0, add)               # lpc_extrapolate_fwd, :288
        rate.push(buf[add + inbuf:add + inbuf + add])
        cut_front = False                                     # on_chunk has done it (:189-195)
    rate.drain()
    tail = np.zeros((0, nch), np.float32)
    while True:
        y = pull_once()
        if y.shape[0] == 0:
            break
        cut = min(drop - dropped, y.shape[0]) if cut_front else 0
        dropped += cut
        tail = np.concatenate([tail, y[cut:]], 0)
        avail = tail.shape[0] - min(tail.shape[0], drop)      # hold back the last `drop` frames
        if avail:
            out.append(tail[:avail])
            tail = tail[avail:]
    return cat()
