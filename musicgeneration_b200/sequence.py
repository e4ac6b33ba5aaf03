"""Events -> notes -> MIDI file: the host-side tail of ``generate.py``.

Mirrors the decode half of MT/sequence.py (``EventSeq.from_array`` :193-205, ``EventSeq.to_note_seq``
:250-284, the constants :7-24) and ``utils.event_indeces_to_midi_file`` (MT/utils.py:25-31).  Pure host
code by nature (a sequential state machine over a few thousand event ids after sampling has finished);
no device work and no pretty_midi: the Standard MIDI File is written here directly (format 1, the
resolution / tempo / program of MT/sequence.py:7-10,65-71).

Vocabulary (MT/sequence.py:207-226): [0, 88) note_on pitch 21.., [88, 176) note_off, [176, 208)
velocity bin, [208, 308) time shift of (k + 1) * 10 ms; ids outside [0, 308) (pad 388, eos ...) are
skipped exactly as the reference's range scan skips them.
"""
from __future__ import annotations

import struct
from typing import List, NamedTuple, Sequence

import numpy as np

DEFAULT_SAVING_PROGRAM = 1
DEFAULT_RESOLUTION = 220
DEFAULT_TEMPO = 120
DEFAULT_VELOCITY = 64
PITCH_START, N_PITCH = 21, 88
VELOCITY_START, VELOCITY_STOP, VELOCITY_STEPS = 21, 109, 32
BEAT_LENGTH = 60 / DEFAULT_TEMPO
TIME_SHIFT_BINS = 0.01 * np.arange(1, 101)
DEFAULT_NOTE_LENGTH = BEAT_LENGTH * 2
MIN_NOTE_LENGTH = BEAT_LENGTH / 2
NOTE_ON, NOTE_OFF, VELOCITY, TIME_SHIFT = 0, N_PITCH, 2 * N_PITCH, 2 * N_PITCH + VELOCITY_STEPS
EVENT_DIM = TIME_SHIFT + TIME_SHIFT_BINS.size


class Note(NamedTuple):
    velocity: int
    pitch: int
    start: float
    end: float


def velocity_bins() -> np.ndarray:
    """MT/sequence.py:228-233."""
    n = VELOCITY_STOP - VELOCITY_START
    return np.arange(VELOCITY_START, VELOCITY_STOP, n / (VELOCITY_STEPS - 1))


def events_to_notes(event_indeces: Sequence[int], velocity_scale: float = None) -> List[Note]:
    """Event ids -> notes sorted by start (stable), as ``EventSeq.from_array(ids).to_note_seq()`` followed
    by ``NoteSeq.__init__``'s filter (end >= start) and sort; ``velocity_scale`` applies the rescaling of
    MT/utils.py:28-29."""
    vbins = velocity_bins()
    time = 0.0
    velocity = DEFAULT_VELOCITY
    notes = []                      # [velocity, pitch, start, end or None]
    open_notes = {}
    for idx in event_indeces:
        idx = int(idx)
        if NOTE_ON <= idx < NOTE_OFF:
            pitch = idx - NOTE_ON + PITCH_START
            note = [velocity, pitch, time, None]
            notes.append(note)
            open_notes[pitch] = note
        elif NOTE_OFF <= idx < VELOCITY:
            pitch = idx - NOTE_OFF + PITCH_START
            note = open_notes.pop(pitch, None)
            if note is not None:
                note[3] = max(time, note[2] + MIN_NOTE_LENGTH)
        elif VELOCITY <= idx < TIME_SHIFT:
            velocity = vbins[min(idx - VELOCITY, vbins.size - 1)]
        elif TIME_SHIFT <= idx < EVENT_DIM:
            time += TIME_SHIFT_BINS[idx - TIME_SHIFT]
    out = []
    for v, pitch, start, end in notes:
        if end is None:
            end = start + DEFAULT_NOTE_LENGTH
        v = int(v)
        if velocity_scale is not None:
            v = int((v - 64) * velocity_scale + 64)
        if end >= start:
            out.append(Note(v, pitch, float(start), float(end)))
    out.sort(key=lambda n: n.start)
    return out


def _vlq(n: int) -> bytes:
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


def notes_to_midi_bytes(notes: Sequence[Note], program: int = DEFAULT_SAVING_PROGRAM,
                        resolution: int = DEFAULT_RESOLUTION, tempo: float = DEFAULT_TEMPO) -> bytes:
    """Standard MIDI File, format 1: a conductor track (tempo, 4/4) and one instrument track named
    'NoteSeq' (MT/sequence.py:65-71); tick = round(seconds * resolution * tempo / 60)."""
    tick = lambda t: int(round(t * resolution * tempo / 60.0))
    t0 = b"\x00\xff\x51\x03" + struct.pack(">I", int(round(60_000_000 / tempo)))[1:]
    t0 += b"\x00\xff\x58\x04\x04\x02\x18\x08" + b"\x00\xff\x2f\x00"
    evs = []
    for order, n in enumerate(notes):
        vel = max(0, min(127, int(n.velocity)))
        evs.append((tick(n.start), 1, order, bytes([0x90, n.pitch & 0x7F, vel])))
        evs.append((tick(n.end), 0, order, bytes([0x90, n.pitch & 0x7F, 0])))       # note-off as velocity-0 note-on
    evs.sort(key=lambda e: (e[0], e[1], e[2]))                                       # offs before ons at equal ticks
    name = b"NoteSeq"
    t1 = b"\x00\xff\x03" + _vlq(len(name)) + name + b"\x00" + bytes([0xC0, program & 0x7F])
    last = 0
    for t, _, _, msg in evs:
        t1 += _vlq(t - last) + msg
        last = t
    t1 += b"\x01\xff\x2f\x00"
    chunk = lambda body: b"MTrk" + struct.pack(">I", len(body)) + body
    return b"MThd" + struct.pack(">IHHH", 6, 1, 2, resolution) + chunk(t0) + chunk(t1)


def event_indeces_to_midi_file(event_indeces, midi_file_name, velocity_scale=0.8):
    """MT/utils.py:25-31 (same name, arguments and return value: the number of notes written)."""
    notes = events_to_notes(event_indeces, velocity_scale)
    with open(midi_file_name, "wb") as f:
        f.write(notes_to_midi_bytes(notes))
    return len(notes)


decode_midi = event_indeces_to_midi_file
