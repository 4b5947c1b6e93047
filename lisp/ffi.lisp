;;; ffi.lisp — CFFI bindings of include/threebz_cuda.h, one DEFCFUN per entry point the shim uses.
(in-package #:3bz)

(cffi:define-foreign-library libthreebz-cuda
  (t (:default "libthreebz_cuda")))
(cffi:use-foreign-library libthreebz-cuda)

(defconstant +tbz-deflate+ 0)
(defconstant +tbz-zlib+ 1)
(defconstant +tbz-gzip+ 2)
(defconstant +tbz-finished+ 0)
(defconstant +tbz-input-underrun+ 1)
(defconstant +tbz-output-overflow+ 2)
(defconstant +tbz-e-buffer-switch+ -5)
(defconstant +tbz-e-state+ -6)

(cffi:defcstruct tbz-member
  (in :pointer) (in-len :uint64) (out :pointer) (out-cap :uint64))
(cffi:defcstruct tbz-result
  (out-len :uint64) (in-used :uint64) (checksum :uint32) (verdict :int32) (where :uint32) (path :uint32))

(cffi:defcfun "tbz_ctx_create" :int32 (device :int32) (flags :uint64) (ctx :pointer))
(cffi:defcfun "tbz_ctx_destroy" :int32 (ctx :pointer))
(cffi:defcfun "tbz_strerror" :string (status :int32))
(cffi:defcfun "tbz_verdict_name" :string (verdict :int32))
(cffi:defcfun "tbz_ctx_last_error" :string (ctx :pointer))
(cffi:defcfun "tbz_host_register" :int32 (p :pointer) (n :uint64) (flags :uint32))
(cffi:defcfun "tbz_host_unregister" :int32 (p :pointer))
(cffi:defcfun "tbz_inflate_batch" :int32
  (ctx :pointer) (format :int32) (members :pointer) (n :uint64) (results :pointer)
  (flags :uint32) (device-ms :pointer))
(cffi:defcfun "tbz_inflate_batch_multi" :int32
  (ctxs :pointer) (g :int32) (format :int32) (members :pointer) (n :uint64) (results :pointer)
  (flags :uint32) (device-ms-per-gpu :pointer))
(cffi:defcfun "tbz_device_count" :int32 (n :pointer))
(cffi:defcfun "tbz_host_alloc" :int32 (n :uint64) (p :pointer))
(cffi:defcfun "tbz_host_free" :int32 (p :pointer))
(cffi:defcfun "tbz_inflate_single" :int32
  (ctx :pointer) (format :int32) (in :pointer) (in-len :uint64) (out :pointer) (out-cap :uint64)
  (result :pointer) (flags :uint32) (device-ms :pointer))
(cffi:defcfun "tbz_inflate_alloc" :int32
  (ctx :pointer) (format :int32) (in :pointer) (in-len :uint64) (out :pointer) (result :pointer))
(cffi:defcfun "tbz_free" :void (p :pointer))
(cffi:defcfun "tbz_session_create" :int32 (ctx :pointer) (format :int32) (session :pointer))
(cffi:defcfun "tbz_session_destroy" :int32 (session :pointer))
(cffi:defcfun "tbz_session_set_output" :int32 (session :pointer) (out :pointer) (cap :uint64))
(cffi:defcfun "tbz_session_rebind_output" :int32 (session :pointer) (out :pointer))
(cffi:defcfun "tbz_session_replace_output" :int32 (session :pointer) (out :pointer) (cap :uint64))
(cffi:defcfun "tbz_session_decompress" :int32
  (session :pointer) (in :pointer) (n :uint64) (ret :pointer) (verdict :pointer))
(cffi:defcfun "tbz_session_flags" :int32
  (session :pointer) (finished :pointer) (underrun :pointer) (overflow :pointer))
(cffi:defcfun "tbz_session_consumed" :int32 (session :pointer) (n :pointer))

(defvar *device* 0 "CUDA device the engine context of this thread is created on.")
(defvar *ctx* nil "tbz_ctx of the current thread (a ctx is single-owner; bind per thread).")

(defun ctx ()
  (or *ctx*
      (cffi:with-foreign-object (p :pointer)
        (check (tbz-ctx-create *device* 0 p))
        (setf *ctx* (cffi:mem-ref p :pointer)))))

(defun check (status)
  "Engine failures (CUDA error, no device, bad argument) are Lisp errors; there is no CPU fallback."
  (unless (zerop status)
    (error "threebz-cuda: ~a~@[ (~a)~]" (tbz-strerror status)
           (and *ctx* (tbz-ctx-last-error *ctx*))))
  status)

(defun format-code (format)
  (ecase format (:deflate +tbz-deflate+) (:zlib +tbz-zlib+) (:gzip +tbz-gzip+)))

;;; gzip member metadata (gzip.lisp:17-28) and the new walker over concatenated members
(cffi:defcstruct tbz-gzip-header
  (verdict :int32) (flags :uint32) (mtime :uint32) (xfl :uint32) (os :uint32) (header-crc :uint32)
  (extra-off :uint64) (extra-len :uint64) (name-off :uint64) (name-len :uint64)
  (comment-off :uint64) (comment-len :uint64) (header-len :uint64))
(cffi:defcfun "tbz_gzip_header_parse" :int32 (in :pointer) (in-len :uint64) (h :pointer))
(cffi:defcfun "tbz_inflate_gzip_members" :int32
  (ctx :pointer) (in :pointer) (in-len :uint64) (out :pointer) (out-cap :uint64) (results :pointer)
  (max-members :uint64) (n-members :pointer) (in-used :pointer))
