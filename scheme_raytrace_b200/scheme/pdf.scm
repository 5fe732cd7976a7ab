;; pdf.scm -- pdf constructors select the estimator of the render call (see srt-scene.scm):
;; (make-mixture-pdf (make-hitable-pdf light #f) (make-cosine-pdf #f)) == estimator 1 + lights
(define-module pdf
  (export make-cosine-pdf make-hitable-pdf make-mixture-pdf))
(select-module pdf)
(define (make-cosine-pdf w) (vector 'pdf 'cosine w #f))
(define (make-hitable-pdf obj origin) (vector 'pdf 'hitable obj origin))
(define (make-mixture-pdf p0 p1) (vector 'pdf 'mixture p0 p1))
