"""CPU: the wire protocol (rust/protocol.md) and the connection front-end over a fake frame service."""

import numpy as np
import pytest

from moshi_b200 import protocol as P
from moshi_b200.serving import ACTIVE, NODATA, RESET


def test_message_framing_round_trips():
    cases = [P.Message(P.MT_HANDSHAKE, (0, 7)), P.Message(P.MT_AUDIO, b"OggS\x00\x01"), P.Message(P.MT_TEXT, " héllo"),
             P.Message(P.MT_CONTROL, P.CONTROL_PAUSE), P.Message(P.MT_METADATA, {"text_temperature": 0.7}),
             P.Message(P.MT_ERROR, "no free slot"), P.Message(P.MT_PING)]
    for m in cases:
        wire = P.encode_message(m)
        assert wire[0] == m.kind
        back = P.decode_message(wire)
        assert back.kind == m.kind and (tuple(back.payload) if m.kind == P.MT_HANDSHAKE else back.payload) == m.payload
    assert P.encode_message(P.Message(P.MT_HANDSHAKE, (0, 1))) == b"\x00" + (0).to_bytes(4, "little") + (1).to_bytes(4, "little")
    assert P.decode_message(b"") is None and P.decode_message(b"\x09junk") is None       # unknown kinds are discarded
    with pytest.raises(ValueError):
        P.decode_message(b"\x03\x07")
    with pytest.raises(ValueError):
        P.encode_message(P.Message(P.MT_CONTROL, 9))


class FakeService:
    """Echoes: PCM out = PCM in * 0.5, text token = frame counter of the slot, ready after one warm-up frame per slot."""

    def __init__(self, B):
        self.seen = np.zeros(B, dtype=np.int64)
        self.calls = []

    def step(self, batch_pcm, pcm_out, tokens_out, updates=None, flags_out=None, noise=None):
        B = pcm_out.shape[0]
        self.calls.append(np.array(updates).copy())
        pcm = batch_pcm.reshape(B, -1)
        for b in range(B):
            if updates[b] == RESET:
                self.seen[b] = 0
            if updates[b] == NODATA:
                flags_out[b] = 0
                continue
            self.seen[b] += 1
            flags_out[b] = 1 if self.seen[b] > 1 else 0
            pcm_out[b] = 0.5 * pcm[b]
            tokens_out[b] = -2 if not flags_out[b] else self.seen[b] + 1        # 3 = padding on the first ready frame


class FakeTokenizer:
    def id_to_piece(self, i):
        return f"▁tok{i}"


def test_front_end_drives_slots_frames_and_outboxes():
    B, fs = 3, 1920
    svc = FakeService(B)
    front = P.FrontEnd(svc, B, fs, text_tokenizer=FakeTokenizer(), model_version=5, codec_factory=lambda: (P.RawPcmCodec(), P.RawPcmCodec()))
    a, b = front.connect(), front.connect()
    assert P.decode_message(a.outbox.pop(0)).payload == (0, 5)          # handshake first (server.py:167)
    b.outbox.clear()
    rng = np.random.default_rng(0)
    pcm_a = rng.standard_normal(3 * fs + 700).astype(np.float32)
    # audio arrives in arbitrary chunks, split across messages, even inside a sample
    wire = pcm_a.astype("<f4").tobytes()
    for lo in range(0, len(wire), 5000):
        front.receive(a, bytes([P.MT_AUDIO]) + wire[lo:lo + 5000])
    front.receive(b, bytes([P.MT_AUDIO]) + rng.standard_normal(fs).astype("<f4").tobytes())
    front.receive(b, b"\x06")                                          # ping: ignored
    front.receive(b, b"")                                              # empty: ignored
    assert front.step() == 0                                           # warm-up frame of both slots: nothing to send yet
    assert list(svc.calls[-1]) == [RESET, RESET, NODATA]
    assert front.step() == 1 and front.step() == 1 and front.step() == 0      # a has two more frames, b none; then nothing buffered
    assert list(svc.calls[-1]) == [ACTIVE, NODATA, NODATA] and len(svc.calls) == 3
    kinds = [m[0] for m in a.outbox]
    assert kinds == [P.MT_AUDIO, P.MT_AUDIO, P.MT_TEXT]                # first ready frame carries text token 3 (padding): no text
    got = np.frombuffer(a.outbox[0][1:], dtype="<f4")
    np.testing.assert_array_equal(got, 0.5 * pcm_a[fs:2 * fs])
    assert P.decode_message(a.outbox[2]).payload == " tok4"
    assert not b.outbox
    # restart: the slot is recycled and its next frame carries RESET
    front.receive(a, bytes([P.MT_CONTROL, P.CONTROL_RESTART]))
    front.receive(a, bytes([P.MT_AUDIO]) + np.zeros(fs, "<f4").tobytes())
    front.step()
    assert svc.calls[-1][a.slot] == RESET
    front.disconnect(a)
    front.disconnect(b)
    assert front.pool.free_slots == B
    c = [front.connect() for _ in range(B)]
    with pytest.raises(RuntimeError):
        front.connect()
    assert len({x.slot for x in c}) == B
