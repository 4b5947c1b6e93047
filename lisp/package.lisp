;;; package.lisp — same package name and export list as the reference (package.lisp:1-27),
;;; plus the batch entry point and the condition type the shim signals.
(defpackage 3bz
  (:use :cl)
  (:import-from :alexandria #:with-gensyms #:once-only)
  (:export
   #:decompress
   #:decompress-vector
   #:with-octet-pointer
   #:make-octet-vector-context
   #:make-octet-stream-context
   #:make-octet-pointer-context
   #:make-deflate-state
   #:make-zlib-state
   #:make-gzip-state
   #:finished
   #:input-underrun
   #:output-overflow
   #:%resync-file-stream
   #:replace-output-buffer
   ;; new
   #:decompress-batch
   #:gzip-header
   #:decompress-gzip-members
   #:*device*))
