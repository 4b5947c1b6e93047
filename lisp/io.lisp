;;; io.lisp — the input contexts of the reference (io-common.lisp:36-45, io-mmap.lisp:21-54) with
;;; the same constructors.  A context no longer feeds a Lisp bit reader word by word
;;; (io.lisp:17-58, io-mmap.lisp:67-114): it names the octets [offset, end) the engine may consume.
(in-package #:3bz)

(deftype octet () '(unsigned-byte 8))
(deftype octet-vector () '(simple-array octet (*)))

(defstruct (context-boxes (:conc-name cb-))
  (start 0 :type fixnum) (end 0 :type fixnum) (offset 0 :type fixnum))

(defclass octet-vector-context ()
  ((octet-vector :reader octet-vector :initarg :octet-vector)
   (boxes :reader boxes :initarg :boxes)))

(defun make-octet-vector-context (vector &key (start 0) (offset start) (end (length vector)))
  (make-instance 'octet-vector-context
                 :octet-vector vector
                 :boxes (make-context-boxes :start start :offset offset :end end)))

;; io-mmap.lisp:21-40: an octet-pointer is valid only inside WITH-OCTET-POINTER.  Here the scope
;; also registers (pins) the range with the CUDA driver, so the engine DMAs straight from a
;; caller's mmap()ed file; if registration fails (file-backed mappings sometimes do) the engine
;; stages through its own pinned buffer instead.
(defclass octet-pointer ()
  ((base :reader base :initarg :base)
   (size :reader size :initarg :size)
   (scope :reader scope :initarg :scope)))

(defmacro with-octet-pointer ((var pointer size) &body body)
  (with-gensyms (scope registered)
    (once-only (pointer size)
      `(let* ((,scope (cons t ',var))
              (,registered (and (plusp ,size) (zerop (tbz-host-register ,pointer ,size 0)))))
         (unwind-protect
              (let ((,var (make-instance 'octet-pointer :base ,pointer :size ,size :scope ,scope)))
                ,@body)
           (setf (car ,scope) nil)
           (when ,registered (tbz-host-unregister ,pointer)))))))

(defun valid-octet-pointer (op)
  (and (car (scope op)) (not (cffi:null-pointer-p (base op))) (plusp (size op))))

(defclass octet-pointer-context ()
  ((op :reader op :initarg :op)
   (pointer :reader %pointer :initarg :pointer)
   (boxes :reader boxes :initarg :boxes)))

(defun make-octet-pointer-context (octet-pointer &key (start 0) (offset 0) (end (size octet-pointer)))
  (make-instance 'octet-pointer-context
                 :op octet-pointer :pointer (base octet-pointer)
                 :boxes (make-context-boxes :start start :offset offset :end end)))

;; octet-stream-context (io-common.lisp:47-63, io.lisp:61-104).  The reference pulls 4 / 8 octets at a
;; time through FILE-POSITION and READ-BYTE (README.md:13 "very slow"); here the unread octets
;; [offset, end) are read with one READ-SEQUENCE per DECOMPRESS call and travel like an octet vector.
(defclass octet-stream-context ()
  ((octet-stream :reader octet-stream :initarg :octet-stream)
   (boxes :reader boxes :initarg :boxes)))

(defun valid-octet-stream (os)
  (and (typep os 'stream) (subtypep (stream-element-type os) 'octet) (open-stream-p os) (input-stream-p os)))

(defun make-octet-stream-context (file-stream &key (start 0) (offset 0) (end (file-length file-stream)))
  (make-instance 'octet-stream-context :octet-stream file-stream
                 :boxes (make-context-boxes :start start :offset offset :end end)))

(defgeneric %resync-file-stream (context)
  (:method (context) (declare (ignore context)) nil)
  (:method ((context octet-stream-context))
    (file-position (octet-stream context) (cb-offset (boxes context)))))

(defgeneric call-with-unread-octets (context function)
  (:documentation "Calls FUNCTION with a foreign pointer to the unread octets and their count."))
(defmethod call-with-unread-octets ((c octet-vector-context) function)
  (let* ((b (boxes c)) (n (- (cb-end b) (cb-offset b))))
    (cffi:with-pointer-to-vector-data (p (octet-vector c))     ; pins the vector for the call
      (funcall function (cffi:inc-pointer p (cb-offset b)) n))))
(defmethod call-with-unread-octets ((c octet-pointer-context) function)
  (unless (valid-octet-pointer (op c))
    (error "trying to use octet-pointer outside scope of with-octet-pointer"))
  (let ((b (boxes c)))
    (funcall function (cffi:inc-pointer (%pointer c) (cb-offset b)) (- (cb-end b) (cb-offset b)))))

(defmethod call-with-unread-octets ((c octet-stream-context) function)
  (let* ((b (boxes c)) (n (max 0 (- (cb-end b) (cb-offset b))))
         (s (octet-stream c))
         (v (make-array n :element-type 'octet)))
    (assert (valid-octet-stream s))
    (file-position s (cb-offset b))
    (let ((got (read-sequence v s)))
      (cffi:with-pointer-to-vector-data (p v)
        (funcall function p got)))))
