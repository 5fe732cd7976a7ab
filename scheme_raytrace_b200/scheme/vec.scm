;; vec.scm -- vec3 helpers for scene scripts (host side only; the device uses fp32 registers).
;; Same exported names as the reference's vec module so scripts that (use vec :prefix v:) load.
(define-module vec
  (use gauche.uvector)
  (export vec3 x y z sum diff prod scale quot dot length sq-len unit cross))
(select-module vec)

(define (vec3 a b c) (f64vector a b c))
(define (x v) (f64vector-ref v 0))
(define (y v) (f64vector-ref v 1))
(define (z v) (f64vector-ref v 2))
(define (fold-vec op first rest)
  (if (null? rest) first (fold-vec op (op first (car rest)) (cdr rest))))
(define (sum . vs) (fold-vec f64vector-add (car vs) (cdr vs)))
(define (diff . vs) (fold-vec f64vector-sub (car vs) (cdr vs)))
;; the reference folds prod from the RIGHT (vec.scm:35-39 reduce-right): v1 * (v2 * v3) - kept, it differs from the left fold by an ulp
(define (prod . vs)
  (let ((r (reverse vs)))
    (fold-vec (lambda (acc v) (f64vector-mul v acc)) (car r) (cdr r))))
(define (quot . vs) (fold-vec f64vector-div (car vs) (cdr vs)))
(define (scale v k) (f64vector-mul v k))
(define (dot a b) (f64vector-dot a b))
(define (sq-len v) (dot v v))
(define (length v) (sqrt (sq-len v)))
(define (unit v) (scale v (/ 1 (length v))))
(define (cross a b)
  (vec3 (- (* (y a) (z b)) (* (y b) (z a)))
        (- (* (z a) (x b)) (* (z b) (x a)))
        (- (* (x a) (y b)) (* (x b) (y a)))))
