;;; api.lisp — decompress-vector, decompress, replace-output-buffer and the status readers with the
;;; signatures of the reference (api.lisp:3-72), plus DECOMPRESS-BATCH.
(in-package #:3bz)

;;; states: (make-deflate-state :output-buffer b) etc. keep their keyword constructor
;;; (deflate.lisp:4, zlib.lisp:3, gzip.lisp:3).  A state owns one tbz_session: the device keeps the
;;; decoded member, DECOMPRESS hands out slices of it.
(defclass deflate-state ()
  ((session :accessor %session :initform nil)
   (format :reader %format :initform :deflate :allocation :class)
   (output-buffer :accessor ds-output-buffer :initarg :output-buffer :initform nil)
   (output-offset :accessor ds-output-offset :initform 0)
   (bound :accessor %bound :initform nil)))
(defclass zlib-state (deflate-state) ((format :initform :zlib :allocation :class)))
(defclass gzip-state (deflate-state) ((format :initform :gzip :allocation :class)))

(defun %make-state (class output-buffer)
  (let ((s (make-instance class :output-buffer output-buffer)))
    (cffi:with-foreign-object (p :pointer)
      (check (tbz-session-create (ctx) (format-code (%format s)) p))
      (setf (%session s) (cffi:mem-ref p :pointer)))
    #+sbcl (let ((h (%session s))) (sb-ext:finalize s (lambda () (tbz-session-destroy h))))
    s))
(defun make-deflate-state (&key output-buffer) (%make-state 'deflate-state output-buffer))
(defun make-zlib-state (&key output-buffer) (%make-state 'zlib-state output-buffer))
(defun make-gzip-state (&key output-buffer) (%make-state 'gzip-state output-buffer))

(defun %flags (state)
  (cffi:with-foreign-objects ((f :int32) (u :int32) (o :int32))
    (check (tbz-session-flags (%session state) f u o))
    (values (plusp (cffi:mem-ref f :int32)) (plusp (cffi:mem-ref u :int32)) (plusp (cffi:mem-ref o :int32)))))
(defun finished (state) (nth-value 0 (%flags state)))
(defun input-underrun (state) (nth-value 1 (%flags state)))
(defun output-overflow (state) (nth-value 2 (%flags state)))

(defun decompress (context state)
  "api.lisp:3-10.  Returns the current offset into the output buffer; sets exactly the three flags."
  (let ((out (or (ds-output-buffer state) (setf (ds-output-buffer state) (make-array 0 :element-type 'octet)))))
    (cffi:with-pointer-to-vector-data (po out)
      ;; SBCL pins OUT only for this call: (re)bind its current address every time
      (if (and (zerop (ds-output-offset state)) (not (%bound state)))
          (progn (check (tbz-session-set-output (%session state) po (length out)))
                 (setf (%bound state) t))
          (check (tbz-session-rebind-output (%session state) po)))
      (call-with-unread-octets
       context
       (lambda (pin n)
         (cffi:with-foreign-objects ((ret :int64) (verdict :int32))
           (let ((rc (tbz-session-decompress (%session state) pin n ret verdict)))
             ;; the session owns every unread octet -- unless the stream finished inside them: the context
             ;; then stops just past the consumed octets (io.lisp:17-58), where %resync-file-stream seeks
             (let ((b (boxes context)))
               (if (zerop rc)
                   (cffi:with-foreign-object (used :uint64)
                     (tbz-session-consumed (%session state) used)
                     (setf (cb-offset b) (min (cb-end b) (+ (cb-offset b) (cffi:mem-ref used :uint64)))))
                   (setf (cb-offset b) (cb-end b))))
             (when (= rc +tbz-e-state+) (error "decompress called on a finished or failed state"))
             (check rc)
             (let ((v (cffi:mem-ref verdict :int32)))
               (when (>= v 16) (error "~a" (tbz-verdict-name v))))    ; where the reference signals
             (setf (ds-output-offset state) (cffi:mem-ref ret :int64)))))))))

(defun replace-output-buffer (state buffer)
  "api.lisp:12-21."
  (cffi:with-pointer-to-vector-data (p buffer)
    (let ((rc (tbz-session-replace-output (%session state) p (length buffer))))
      (when (= rc +tbz-e-buffer-switch+)
        (error "can't switch buffers without filling old one yet."))
      (check rc)))
  (setf (ds-output-buffer state) buffer
        (ds-output-offset state) 0
        (%bound state) t))

(defun %verdict-error (verdict format)
  (cond ((= verdict +tbz-input-underrun+) (error "incomplete ~a stream" format))
        ((= verdict +tbz-output-overflow+) (error "not enough space to decompress ~a stream" format))
        (t (error "~a" (tbz-verdict-name verdict)))))

(defun decompress-vector (compressed &key (format :zlib) (start 0) (end (length compressed)) output)
  "api.lisp:23-65: returns (values buffer count)."
  (cffi:with-foreign-object (r '(:struct tbz-result))
    (cffi:with-pointer-to-vector-data (pin compressed)
      (if output
          (cffi:with-pointer-to-vector-data (pout output)
            (check (tbz-inflate-single (ctx) (format-code format) (cffi:inc-pointer pin start) (- end start)
                                       pout (length output) r 0 (cffi:null-pointer)))
            (let ((v (cffi:foreign-slot-value r '(:struct tbz-result) 'verdict)))
              (unless (= v +tbz-finished+) (%verdict-error v format)))
            (values output (cffi:foreign-slot-value r '(:struct tbz-result) 'out-len)))
          ;; no :output — the reference grows 32 KiB buffers by doubling and concatenates
          ;; (api.lisp:50-65); the engine sizes the result on the device and returns it whole
          (cffi:with-foreign-object (pp :pointer)
            (check (tbz-inflate-alloc (ctx) (format-code format) (cffi:inc-pointer pin start) (- end start) pp r))
            (let* ((p (cffi:mem-ref pp :pointer))
                   (v (cffi:foreign-slot-value r '(:struct tbz-result) 'verdict))
                   (n (cffi:foreign-slot-value r '(:struct tbz-result) 'out-len)))
              (unwind-protect
                   (progn
                     (unless (= v +tbz-finished+) (%verdict-error v format))
                     (let ((b (make-array n :element-type 'octet)))
                       (cffi:with-pointer-to-vector-data (pb b)
                         (cffi:foreign-funcall "memcpy" :pointer pb :pointer p :size n :pointer))
                       (values b n)))
                (tbz-free p))))))))

;;; one engine context per device for the multi-GPU batch entry point
(defvar *device-ctxs* (make-hash-table) "device index -> tbz_ctx of this thread")
(defun device-ctx (device)
  (or (gethash device *device-ctxs*)
      (setf (gethash device *device-ctxs*)
            (if (eql device *device*)
                (ctx)
                (cffi:with-foreign-object (c :pointer)
                  (check (tbz-ctx-create device 0 c))
                  (cffi:mem-ref c :pointer))))))

(defun decompress-batch (members &key (format :zlib) capacities devices)
  "NEW: many independent members in one engine call.  MEMBERS: sequence of octet-vectors;
CAPACITIES: per-member output size (one integer or a sequence); DEVICES: list of CUDA device indices to
shard the members over (host-side partition, tbz_inflate_batch_multi) -- default: this thread's device.
Returns a list of (buffer count verdict) -- verdict :finished / :input-underrun / :output-overflow or the
engine's name for the place where the reference would have signalled.  A bad member never poisons the batch.
The members travel through two pinned arenas (inputs back to back, outputs back to back): the engine then
DMAs straight from and to them, and only one Lisp vector is pinned at a time however long the batch is."
  (let* ((ins (coerce members 'list))
         (n (length ins))
         (caps (if (integerp capacities) (make-list n :initial-element capacities) (coerce capacities 'list)))
         (outs (mapcar (lambda (c) (make-array c :element-type 'octet)) caps))
         (in-total (reduce #'+ ins :key #'length))
         (out-total (reduce #'+ caps)))
    (cffi:with-foreign-objects ((m '(:struct tbz-member) (max 1 n)) (r '(:struct tbz-result) (max 1 n))
                                (ph-in :pointer) (ph-out :pointer))
      (check (tbz-host-alloc (max 1 in-total) ph-in))
      (let ((h-in (cffi:mem-ref ph-in :pointer)) (h-out (cffi:null-pointer)))
        (unwind-protect
             (progn
               (check (tbz-host-alloc (max 1 out-total) ph-out))
               (setf h-out (cffi:mem-ref ph-out :pointer))
               (loop with io = 0 and oo = 0
                     for i from 0 for v in ins for c in caps
                     for e = (cffi:mem-aptr m '(:struct tbz-member) i)
                     do (cffi:with-pointer-to-vector-data (p-in v)
                          (cffi:foreign-funcall "memcpy" :pointer (cffi:inc-pointer h-in io) :pointer p-in
                                                         :size (length v) :pointer))
                        (setf (cffi:foreign-slot-value e '(:struct tbz-member) 'in) (cffi:inc-pointer h-in io)
                              (cffi:foreign-slot-value e '(:struct tbz-member) 'in-len) (length v)
                              (cffi:foreign-slot-value e '(:struct tbz-member) 'out) (cffi:inc-pointer h-out oo)
                              (cffi:foreign-slot-value e '(:struct tbz-member) 'out-cap) c)
                        (incf io (length v)) (incf oo c))
               (if (and devices (rest devices))
                   (let ((g (length devices)))
                     (cffi:with-foreign-object (cs :pointer g)
                       (loop for d in devices for k from 0
                             do (setf (cffi:mem-aref cs :pointer k) (device-ctx d)))
                       (check (tbz-inflate-batch-multi cs g (format-code format) m n r 0 (cffi:null-pointer)))))
                   (check (tbz-inflate-batch (if devices (device-ctx (first devices)) (ctx))
                                             (format-code format) m n r 0 (cffi:null-pointer))))
               (loop with oo = 0
                     for i from 0 for o in outs for c in caps
                     for e = (cffi:mem-aptr r '(:struct tbz-result) i)
                     for v = (cffi:foreign-slot-value e '(:struct tbz-result) 'verdict)
                     for got = (min c (cffi:foreign-slot-value e '(:struct tbz-result) 'out-len))
                     do (when (plusp got)
                          (cffi:with-pointer-to-vector-data (p-out o)
                            (cffi:foreign-funcall "memcpy" :pointer p-out :pointer (cffi:inc-pointer h-out oo)
                                                           :size got :pointer)))
                        (incf oo c)
                     collect (list o (cffi:foreign-slot-value e '(:struct tbz-result) 'out-len)
                                   (case v (0 :finished) (1 :input-underrun) (2 :output-overflow)
                                     (t (tbz-verdict-name v))))))
          (tbz-host-free h-in)
          (unless (cffi:null-pointer-p h-out) (tbz-host-free h-out)))))))

(defun gzip-header (octets &key (start 0) (end (length octets)))
  "The slots DECOMPRESS-GZIP fills from a member header in the reference's GZIP-STATE
\(gzip.lisp:17-28, :113-260), as a plist: :flags :extra :name :comment :operating-system
:mtime/unix :mtime/universal :compression-level :header-length.  NIL while the header is
incomplete (the reference's input-underrun); header errors signal like the reference."
  (cffi:with-foreign-object (h '(:struct tbz-gzip-header))
    (cffi:with-pointer-to-vector-data (pin octets)
      (check (tbz-gzip-header-parse (cffi:inc-pointer pin start) (- end start) h)))
    (flet ((f (slot) (cffi:foreign-slot-value h '(:struct tbz-gzip-header) slot)))
      (let ((v (f 'verdict)) (flg (f 'flags)) (mtime (f 'mtime)) (os (f 'os)) (xfl (f 'xfl)))
        (cond
          ((= v +tbz-input-underrun+) nil)
          ((/= v +tbz-finished+) (%verdict-error v :gzip))
          (t
           (flet ((text (off len)
                    (let ((raw (subseq octets (+ start off) (+ start off len))))
                      ;; rfc says 8859-1, but try utf8 anyway (gzip.lisp:214-217)
                      (or (ignore-errors (babel:octets-to-string raw :encoding :utf-8 :errorp t))
                          (babel:octets-to-string raw :encoding :iso-8859-1)))))
             (list :flags (loop for (bit name) in '((1 :text) (2 :header-crc) (4 :extra) (8 :name) (16 :comment))
                                when (logtest bit flg) collect name)
                   :extra (when (logtest 4 flg) (subseq octets (+ start (f 'extra-off)) (+ start (f 'extra-off) (f 'extra-len))))
                   :name (when (logtest 8 flg) (text (f 'name-off) (f 'name-len)))
                   :comment (when (logtest 16 flg) (text (f 'comment-off) (f 'comment-len)))
                   :operating-system (if (<= 0 os 13)
                                         (aref #(:fat :amiga :vms :unix :vm/cms :atari-tos :hpfs :macintosh
                                                 :z-system :cp/m :tops-20 :ntfs :qdos :acorn-riscos) os)
                                         (list :unknown os))
                   :mtime/unix (unless (zerop mtime) mtime)
                   :mtime/universal (unless (zerop mtime) (+ mtime (encode-universal-time 0 0 0 1 1 1970 0)))
                   :compression-level (or (case xfl (2 :maximum) (4 :fastest)) xfl)
                   :header-length (f 'header-len)))))))))

(defun decompress-gzip-members (octets output &key (max-members (ash 1 20)))
  "NEW: the reference stops after the first member of a gzip file (gzip.lisp:279-286); this walks
all concatenated members into OUTPUT.  Returns (values list-of-(count in-used crc32 verdict) octets-consumed)."
  (let ((n (min max-members (1+ (floor (length octets) 18)))))
    (cffi:with-foreign-objects ((r '(:struct tbz-result) n) (nm :uint64) (used :uint64))
      (cffi:with-pointer-to-vector-data (p-in octets)
        (cffi:with-pointer-to-vector-data (p-out output)
          (check (tbz-inflate-gzip-members (ctx) p-in (length octets) p-out (length output) r n nm used))))
      (values (loop for i below (cffi:mem-ref nm :uint64)
                    for e = (cffi:mem-aptr r '(:struct tbz-result) i)
                    collect (list (cffi:foreign-slot-value e '(:struct tbz-result) 'out-len)
                                  (cffi:foreign-slot-value e '(:struct tbz-result) 'in-used)
                                  (cffi:foreign-slot-value e '(:struct tbz-result) 'checksum)
                                  (cffi:foreign-slot-value e '(:struct tbz-result) 'verdict)))
              (cffi:mem-ref used :uint64)))))
