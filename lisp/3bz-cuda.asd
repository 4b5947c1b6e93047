;;; 3bz-cuda.asd — 3bz's public API re-hosted over libthreebz_cuda.so (B200 inflate engine).
;;; Drop-in for the reference system definition (3bz.asd:1-29): same package, same exports, plus
;;; DECOMPRESS-BATCH.  The decode core (deflate.lisp, huffman-tree.lisp, checksums.lisp and the
;;; arithmetic of zlib.lisp / gzip.lisp) lives in the CUDA library; only the boundary is Lisp.
;;; NOTE: no Common Lisp implementation exists in the build image, so these files are exercised
;;; through the Python mirror 3bz_b200/api.py, which issues the same C-ABI call sequences.
(defsystem :3bz-cuda
  :description "deflate/zlib/gzip decompressor: 3bz API over a CUDA (sm_100a) engine"
  :depends-on (alexandria cffi trivial-features babel)
  :serial t
  :license "MIT"
  :components
  ((:file "package")
   (:file "ffi")
   (:file "io")
   (:file "api")))
