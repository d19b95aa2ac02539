"""SampleView — format table and slicing (reference lib/samples.js:15-183, 253-258).

Only the host-side bookkeeping lives here (format aliases, bytes per sample, sampleCount,
slice); decoding itself runs on the GPU (csrc/sp_device.cuh).
"""
import numpy as np

from . import _lib

# format -> (sampleWidth bytes, typed-array element bytes)   lib/samples.js:30-139
_TABLE = {
    "CU4": (1, 1), "CS4": (1, 1), "CU8": (2, 1), "CS8": (2, 1), "CU12": (3, 1), "CS12": (3, 1),
    "CU16": (4, 2), "CS16": (4, 2), "CU32": (8, 4), "CS32": (8, 4), "CU64": (16, 4), "CS64": (16, 4),
    "CF32": (8, 4), "CF64": (16, 8),
}
_ALIASES = {"DATA": "CU8", "COMPLEX16U": "CU8", "COMPLEX16S": "CS8", "CFILE": "CF32", "COMPLEX": "CF32"}
_AUDIO = {"WAV", "BWF", "WEBM", "OGG", "OPUS", "FLAC", "MP4", "M4A", "AAC", "MP3"}   # lib/samples.js:141-148


class SampleView:
    def __init__(self, format, buffer=None, sampleRate=None, centerFreq=None):
        self.sampleRate = sampleRate or 250000
        self.centerFreq = centerFreq or 0
        format = format.upper()
        self.format = format
        canon = _ALIASES.get(format, format)
        self.audio = canon in _AUDIO                      # lib/samples.js:141-148: decoded to interleaved CF32
        if self.audio:
            self.container = canon
            canon = "CF32"
            self.format = "CF32"                         # "force format on decompressed buffer"
        if canon not in _TABLE:
            canon = "CU8"                      # lib/samples.js:149-155: default to CU8
        self.canonical = canon
        self.sampleWidth, self.elementSize = _TABLE[canon]
        self.buffer = None
        self.sampleCount = 0
        if buffer is not None:
            self.loadBuffer(buffer)

    def loadBuffer(self, buffer):
        if self.audio:                                    # readAudio() + interleaved(), lib/samples.js:260-302
            from .ingest import AUDIO_PCM, decode_wav
            if self.container not in AUDIO_PCM:
                raise NotImplementedError("%s needs the browser's decodeAudioData; decode to WAV or CF32 first" % self.container)
            data, rate, _ch = decode_wav(bytes(buffer))
            self.sampleRate = rate
            self.buffer = data.tobytes()
            self.sampleCount = len(data) // 2
            return self
        # zero copy for anything that exposes the buffer protocol: a page-locked numpy array from ingest.load_capture() stays
        # page-locked (slices are views), so Engine.make_request hands the pinned pointer to the pipelined host path
        if isinstance(buffer, np.ndarray):
            buffer = np.ascontiguousarray(buffer).view(np.uint8).reshape(-1)
        elif not isinstance(buffer, (bytes, bytearray, memoryview)):
            buffer = memoryview(buffer).cast("B")
        if len(buffer) % self.elementSize:
            raise ValueError("RangeError: byte length of typed array should be a multiple of %d" % self.elementSize)
        self.buffer = buffer
        self.sampleCount = len(buffer) / self.sampleWidth     # lib/samples.js:167 (may be fractional)
        return self

    @property
    def duration(self):
        return self.sampleCount / self.sampleRate

    def slice(self, sliceIndex, sliceCount, startSample=0, endSample=0):
        """lib/samples.js:253-258"""
        startSample = startSample or 0
        endSample = endSample or int(self.sampleCount)
        sliceLength = self.sampleWidth * int((endSample - startSample) / sliceCount)
        a = startSample * self.sampleWidth + sliceLength * sliceIndex
        return self.buffer[a:a + sliceLength]

    def format_id(self):
        return _lib.FORMATS.index(self.canonical)
