"""Test-only engine over the CPU logic-emulation build of the kernel sources.

`EmuEngine` lets the GPU-less CI container drive the HOST layer (fit driver, model / optimizer
mirrors, sharding) through exactly the same C-ABI calls the product makes, with host pointers and
the emulation library (csrc/libscaml_emu.so, built from the same .cuh files with -DSCAML_EMU).
It lives under tests/ and is never importable from the package: the product `Engine` raises
without CUDA and without libscaml_b200.so.
"""
import torch

from scamlgp_b200.engine import Engine


class EmuEngine(Engine):
    def __init__(self, lib):
        self.device = torch.device("cpu")
        self.lib = lib
        self._ws = self._pws = self._cws = self._tws = self._gws = self._vws = None
        self.launches = 0

    def _stream(self) -> int:
        return 0
