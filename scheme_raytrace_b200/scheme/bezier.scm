;; bezier.scm -- make-bezier (cubic curve with width) and make-bezier-patch (4x4 control net,
;; an extension the reference does not have).
(define-module bezier
  (use vec :prefix v:)
  (use geometry :prefix g:)
  (export make-bezier make-bezier-patch))
(select-module bezier)

(define (v3->list p) (list (v:x p) (v:y p) (v:z p)))
(define (make-bezier a b c d width material)
  (g:make-obj 5 material (append (v3->list a) (v3->list b) (v3->list c) (v3->list d) (list width)) '()))
;; rows: list of 4 lists of 4 vec3, P[i][j] with i along u
(define (make-bezier-patch rows material)
  (g:make-obj 7 material (apply append (map (lambda (row) (apply append (map v3->list row))) rows)) '()))
