;; texture.scm -- texture constructors that record table rows: #(texture kind rgb scale even odd [image])
(define-module texture
  (export constant-texture checker-texture noise-texture marble-texture image-texture
          texture-kind texture-rgb texture-scale texture-even texture-odd texture-image))
(select-module texture)

(define (mk kind rgb scale even odd) (vector 'texture kind rgb scale even odd))
(define (texture-kind t) (vector-ref t 1))
(define (texture-rgb t) (vector-ref t 2))
(define (texture-scale t) (vector-ref t 3))
(define (texture-even t) (vector-ref t 4))
(define (texture-odd t) (vector-ref t 5))
(define (constant-texture color) (mk 0 color 0 #f #f))
(define (checker-texture even-tex odd-tex) (mk 1 #f 0 even-tex odd-tex))
(define (noise-texture sc) (mk 2 #f sc #f #f))
(define (marble-texture sc) (mk 3 #f sc #f #f))
;; image-texture (reference texture.scm:36-50): data = vector of nx*ny*3 numbers 0..255, top row first;
;; recorded as (nx ny data) in slot 6, written to the scene file's images section by srt-scene
(define (texture-image t) (if (> (vector-length t) 6) (vector-ref t 6) #f))
(define (image-texture data nx ny) (vector 'texture 4 #f 0 #f #f (list nx ny data)))
