"""Host-side mirror of the `panSVR fc_aln` entry over the C ABI (include/pansvr_b200.h).

`AlnContext(index_dir, header_sam)` loads the deBGA index of the SV anchor reference into HBM;
`align_fastq(text)` realigns one block of interleaved signal read pairs and returns the SAM body text of the
main output and of the `-p` output, byte-identical to `panSVR fc_aln -t 1 -S` (deCOY_CLASSIFY_MAIN::init_run,
src/PanSVgenerateVCF/read_realignment.cpp:26).  `fc_aln_main(argv)` is the reference's command line.
No CPU fallback: without the CUDA library and a B200 these raise.
"""
from __future__ import annotations

import ctypes as C

from .ksw import load_library


class AlnOptionsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("match", "mismatch", "gap_open", "gap_ex", "gap_open2", "gap_ex2", "zdrop", "band_width",
                                         "not_ori", "max_use_read", "threads", "explicit_mask")]


class AlnPieceC(C.Structure):
    _fields_ = [("fastq", C.c_void_p), ("fastq_bytes", C.c_size_t), ("await_path", C.c_char_p), ("publish_path", C.c_char_p),
                ("sam_bytes", C.c_size_t), ("ori_bytes", C.c_size_t)]


class AlnStatsC(C.Structure):
    _fields_ = [("reads", C.c_int64), ("mems", C.c_int64), ("ksw_tasks", C.c_int64), ("ksw_cells", C.c_int64),
                ("deferred_pairs", C.c_int64), ("stage_seconds", C.c_double * 8),
                ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("seed_probes", C.c_int64),
                ("seed_kernel_ms", C.c_double), ("ksw_kernel_ms", C.c_double), ("stage_kernel_ms", C.c_double),
                ("stage_kernel_ms_by", C.c_double * 8),
                ("in_order_seconds", C.c_double), ("in_order_pairs", C.c_int64), ("in_order_draws", C.c_int64), ("host_pairs", C.c_int64), ("tie_pairs", C.c_int64)]


def _bind(lib):
    lib.pansvr_aln_create.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(AlnOptionsC), C.c_int, C.POINTER(C.c_void_p)]
    lib.pansvr_aln_create_multi.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(AlnOptionsC), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
    lib.pansvr_aln_destroy.argtypes = [C.c_void_p]
    lib.pansvr_aln_header_text.restype = C.c_char_p
    lib.pansvr_aln_header_text.argtypes = [C.c_void_p]
    lib.pansvr_aln_last_error.restype = C.c_char_p
    lib.pansvr_aln_block.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.pansvr_aln_block_bam.argtypes = lib.pansvr_aln_block.argtypes
    lib.pansvr_bam_open.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.pansvr_bam_write.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.pansvr_bam_close.argtypes = [C.c_void_p]
    lib.pansvr_aln_last_stats.argtypes = [C.c_void_p, C.POINTER(AlnStatsC)]
    lib.pansvr_aln_reset.argtypes = [C.c_void_p]
    lib.pansvr_free.argtypes = [C.c_void_p]
    lib.pansvr_fc_aln_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    lib.pansvr_aln_prime_read_stats.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.pansvr_aln_await_state.argtypes = [C.c_void_p, C.c_char_p]
    lib.pansvr_aln_publish_state.argtypes = [C.c_void_p, C.c_char_p]
    lib.pansvr_aln_pieces.argtypes = [C.c_void_p, C.POINTER(AlnPieceC), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    return lib


class AlnContext:
    def __init__(self, index_dir: str, header_sam: str, device: int = 0, lib=None, **options):
        self.lib = _bind(lib or load_library())
        o = AlnOptionsC(**options)
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):                     # several GPUs of one box
            arr = (C.c_int * len(device))(*device)
            rc = self.lib.pansvr_aln_create_multi(index_dir.encode(), header_sam.encode(), C.byref(o), arr, len(device), C.byref(h))
        else:
            rc = self.lib.pansvr_aln_create(index_dir.encode(), header_sam.encode(), C.byref(o), device, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"pansvr_aln_create failed ({rc}): {self.lib.pansvr_aln_last_error().decode()}")
        self.h = h

    def header_text(self) -> str:
        return self.lib.pansvr_aln_header_text(self.h).decode()

    def align_fastq(self, fastq: bytes):
        s, o = C.c_void_p(), C.c_void_p()
        sl, ol = C.c_size_t(), C.c_size_t()
        rc = self.lib.pansvr_aln_block(self.h, fastq, len(fastq), C.byref(s), C.byref(sl), C.byref(o), C.byref(ol))
        if rc != 0:
            raise RuntimeError(f"pansvr_aln_block failed ({rc}): {self.lib.pansvr_aln_last_error().decode()}")
        try:
            return C.string_at(s, sl.value), C.string_at(o, ol.value)
        finally:
            self.lib.pansvr_free(s); self.lib.pansvr_free(o)

    def align_fastq_view(self, fastq: bytes):
        """Like align_fastq without the copy into Python bytes: returns (sam, ori, release) where sam/ori are memoryviews of
        the library's buffers and release() frees them (call it once the views are no longer used)."""
        s, o = C.c_void_p(), C.c_void_p()
        sl, ol = C.c_size_t(), C.c_size_t()
        rc = self.lib.pansvr_aln_block(self.h, fastq, len(fastq), C.byref(s), C.byref(sl), C.byref(o), C.byref(ol))
        if rc != 0:
            raise RuntimeError(f"pansvr_aln_block failed ({rc}): {self.lib.pansvr_aln_last_error().decode()}")
        sam = memoryview((C.c_char * sl.value).from_address(s.value)).cast("B") if sl.value else memoryview(b"")
        ori = memoryview((C.c_char * ol.value).from_address(o.value)).cast("B") if ol.value else memoryview(b"")

        def release():
            sam.release(); ori.release()
            self.lib.pansvr_free(s); self.lib.pansvr_free(o)
        return sam, ori, release

    def align_ptr(self, addr: int, nbytes: int):
        """pansvr_aln_block on a raw host buffer (e.g. pinned memory): returns ((sam_addr, sam_bytes), (ori_addr, ori_bytes), release)."""
        s, o = C.c_void_p(), C.c_void_p()
        sl, ol = C.c_size_t(), C.c_size_t()
        rc = self.lib.pansvr_aln_block(self.h, C.cast(C.c_void_p(addr), C.c_char_p), nbytes, C.byref(s), C.byref(sl), C.byref(o), C.byref(ol))
        if rc != 0:
            raise RuntimeError(f"pansvr_aln_block failed ({rc}): {self.lib.pansvr_aln_last_error().decode()}")

        def release():
            self.lib.pansvr_free(s); self.lib.pansvr_free(o)
        return (s.value, sl.value), (o.value, ol.value), release

    def align_bytes_at(self, fastq: bytes, offset: int, nbytes: int):
        """pansvr_aln_block on fastq[offset : offset + nbytes] without slicing (no copy); same returns as align_ptr."""
        base = C.cast(C.c_char_p(fastq), C.c_void_p).value
        return self.align_ptr(base + offset, nbytes)

    def align_pieces(self, pieces):
        """pansvr_aln_pieces: pieces = [(addr, nbytes, await_path or None, publish_path or None), ...] of one input, in input order.
        Returns ((sam_addr, sam_bytes), (ori_addr, ori_bytes), [(sam_bytes, ori_bytes) per piece], release)."""
        arr = (AlnPieceC * len(pieces))()
        for a, (addr, n, aw, pb) in zip(arr, pieces):
            a.fastq = addr; a.fastq_bytes = n
            a.await_path = aw.encode() if aw else None
            a.publish_path = pb.encode() if pb else None
        s, o = C.c_void_p(), C.c_void_p()
        sl, ol = C.c_size_t(), C.c_size_t()
        rc = self.lib.pansvr_aln_pieces(self.h, arr, len(pieces), C.byref(s), C.byref(sl), C.byref(o), C.byref(ol))
        if rc != 0:
            raise RuntimeError(f"pansvr_aln_pieces failed ({rc}): {self.lib.pansvr_aln_last_error().decode()}")

        def release():
            self.lib.pansvr_free(s); self.lib.pansvr_free(o)
        return (s.value, sl.value), (o.value, ol.value), [(a.sam_bytes, a.ori_bytes) for a in arr], release

    def prime_read_stats(self, fastq_head: bytes) -> None:
        """Show the context the first record of the whole input (STAT_ fields); needed when its own blocks start later in the input."""
        if self.lib.pansvr_aln_prime_read_stats(self.h, fastq_head, len(fastq_head)) != 0:
            raise RuntimeError(self.lib.pansvr_aln_last_error().decode())

    def await_state(self, path: str) -> None:
        if self.lib.pansvr_aln_await_state(self.h, path.encode()) != 0:
            raise RuntimeError(self.lib.pansvr_aln_last_error().decode())

    def publish_state(self, path: str) -> None:
        if self.lib.pansvr_aln_publish_state(self.h, path.encode()) != 0:
            raise RuntimeError(self.lib.pansvr_aln_last_error().decode())

    def align_fastq_bam(self, fastq: bytes):
        """Like align_fastq, records in uncompressed BAM form (what htslib's bam_write1 hands to BGZF)."""
        s, o = C.c_void_p(), C.c_void_p()
        sl, ol = C.c_size_t(), C.c_size_t()
        rc = self.lib.pansvr_aln_block_bam(self.h, fastq, len(fastq), C.byref(s), C.byref(sl), C.byref(o), C.byref(ol))
        if rc != 0:
            raise RuntimeError(f"pansvr_aln_block_bam failed ({rc}): {self.lib.pansvr_aln_last_error().decode()}")
        try:
            return C.string_at(s, sl.value), C.string_at(o, ol.value)
        finally:
            self.lib.pansvr_free(s); self.lib.pansvr_free(o)

    def write_bam(self, path: str, record_chunks) -> None:
        """BAM file = header + the given chunks of records (outputs of align_fastq_bam), BGZF-compressed like htslib does."""
        f = C.c_void_p()
        if self.lib.pansvr_bam_open(self.h, path.encode(), C.byref(f)) != 0:
            raise RuntimeError(self.lib.pansvr_aln_last_error().decode())
        try:
            for chunk in record_chunks:
                if self.lib.pansvr_bam_write(f, chunk, len(chunk)) != 0:
                    raise RuntimeError(self.lib.pansvr_aln_last_error().decode())
        finally:
            if self.lib.pansvr_bam_close(f) != 0:
                raise RuntimeError(self.lib.pansvr_aln_last_error().decode())

    def reset(self):
        self.lib.pansvr_aln_reset(self.h)

    def stats(self) -> dict:
        st = AlnStatsC()
        self.lib.pansvr_aln_last_stats(self.h, C.byref(st))
        d = {k: getattr(st, k) for k in ("reads", "mems", "ksw_tasks", "ksw_cells", "deferred_pairs", "kernel_launches", "h2d_bytes",
                                         "d2h_bytes", "seed_probes", "seed_kernel_ms", "ksw_kernel_ms", "stage_kernel_ms",
                                         "in_order_seconds", "in_order_pairs", "in_order_draws", "host_pairs", "tie_pairs")}
        d["stage_seconds"] = list(st.stage_seconds)
        d["stage_kernel_ms_by"] = list(st.stage_kernel_ms_by)
        return d

    def close(self):
        if self.h:
            self.lib.pansvr_aln_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fc_aln_main(argv, lib=None) -> int:
    """`panSVR fc_aln` command line: fc_aln_main(["-t", "1", "-S", "-o", out, "-p", ori, index_dir, reads_fq, header_sam])."""
    lib = _bind(lib or load_library())
    args = [b"fc_aln"] + [a.encode() if isinstance(a, str) else a for a in argv]
    arr = (C.c_char_p * (len(args) + 1))(*args, None)
    return lib.pansvr_fc_aln_main(len(args), arr)
