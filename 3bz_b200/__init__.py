"""3bz_b200 — B200-native inflate engine behind 3bz's API (host-side mirror of the Lisp package).

Import with importlib.import_module("3bz_b200") (the name starts with a digit, like the reference's
package `3bz`), or through the `threebz_b200` alias module at the repo root.
"""
from .api import (ThreeBzError, Ctx, default_ctx, device_count, decompress, decompress_vector, decompress_batch,  # noqa: F401
                  with_octet_pointer, make_octet_vector_context, make_octet_stream_context,
                  make_octet_pointer_context, make_deflate_state, make_zlib_state, make_gzip_state,
                  finished, input_underrun, output_overflow, replace_output_buffer,
                  gzip_header, decompress_gzip_members, resync_file_stream)
from . import _ffi, shard  # noqa: F401

# package.lisp:13-27 — the exported symbols, plus the new batch entry point
__all__ = ["decompress", "decompress_vector", "with_octet_pointer", "make_octet_vector_context",
           "make_octet_stream_context", "make_octet_pointer_context", "make_deflate_state",
           "make_zlib_state", "make_gzip_state", "finished", "input_underrun", "output_overflow",
           "replace_output_buffer", "decompress_batch", "gzip_header", "decompress_gzip_members", "resync_file_stream"]
