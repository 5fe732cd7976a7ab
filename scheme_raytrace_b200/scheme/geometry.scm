;; geometry.scm -- object constructors with the reference's names and arity; every object is the
;; tagged vector #(object kind material params children).  Hit testing happens on the GPU.
(define-module geometry
  (use vec :prefix v:)
  (use material :prefix m:)
  (export make-scene scene-obj-list scene-num-obj scene-camera scene-sky-function material
          make-sphere make-moving-sphere make-xy-rect make-xz-rect make-yz-rect flip-normals
          make-box translate rotate-y make-bvh-node make-bvh-with-sah make-constant-medium make-klein
          obj-kind obj-params obj-children make-obj))
(select-module geometry)

(define pi/180 (/ 3.141592653589793 180))
(define (make-obj kind mat params children) (vector 'object kind mat params children))
(define (obj-kind o) (vector-ref o 1))
(define (material o) (vector-ref o 2))
(define (obj-params o) (vector-ref o 3))
(define (obj-children o) (vector-ref o 4))
(define (v3->list p) (list (v:x p) (v:y p) (v:z p)))

;; kinds: 0 sphere 1 moving-sphere 2 xy-rect 3 xz-rect 4 yz-rect 5 bezier 6 constant-medium 7 patch 8 klein
;;        16 flip 17 list 18 translate 19 rotate-y   (== include/srt.h and host/geometry.py)
(define (make-scene obj-list camera sky-function) (vector 'scene #f obj-list camera sky-function))
(define (scene-obj-list s) (vector-ref s 2))
(define (scene-num-obj s) (length (scene-obj-list s)))
(define (scene-camera s) (vector-ref s 3))
(define (scene-sky-function s) (vector-ref s 4))

(define (make-sphere center radius mat) (make-obj 0 mat (append (v3->list center) (list radius)) '()))
(define (make-moving-sphere center0 center1 time0 time1 radius mat)
  (make-obj 1 mat (append (v3->list center0) (list radius) (v3->list center1) (list time0 time1)) '()))
(define (make-xy-rect x0 x1 y0 y1 k mat) (make-obj 2 mat (list x0 x1 y0 y1 k) '()))
(define (make-xz-rect x0 x1 z0 z1 k mat) (make-obj 3 mat (list x0 x1 z0 z1 k) '()))
(define (make-yz-rect y0 y1 z0 z1 k mat) (make-obj 4 mat (list y0 y1 z0 z1 k) '()))
(define (flip-normals obj) (make-obj 16 (material obj) '() (list obj)))
(define (make-box p0 p1 mat)
  (make-obj 17 mat '()
            (list (make-xy-rect (v:x p0) (v:x p1) (v:y p0) (v:y p1) (v:z p1) mat)
                  (flip-normals (make-xy-rect (v:x p0) (v:x p1) (v:y p0) (v:y p1) (v:z p0) mat))
                  (make-xz-rect (v:x p0) (v:x p1) (v:z p0) (v:z p1) (v:y p1) mat)
                  (flip-normals (make-xz-rect (v:x p0) (v:x p1) (v:z p0) (v:z p1) (v:y p0) mat))
                  (make-yz-rect (v:y p0) (v:y p1) (v:z p0) (v:z p1) (v:x p1) mat)
                  (flip-normals (make-yz-rect (v:y p0) (v:y p1) (v:z p0) (v:z p1) (v:x p0) mat)))))
(define (translate obj offset) (make-obj 18 #f (v3->list offset) (list obj)))
(define (rotate-y obj angle)
  (let ((radians (* pi/180 angle)))
    (make-obj 19 (material obj) (list (sin radians) (cos radians) angle) (list obj))))
;; the CPU BVH builders become grouping hints: the GPU LBVH is built over the flattened leaves
(define (make-bvh-node obj-list time0 time1) (make-obj 17 #f '() obj-list))
(define (make-bvh-with-sah obj-list time0 time1) (make-obj 17 #f '() obj-list))
(define (make-klein center mat) (make-obj 8 mat (v3->list center) '()))
(define (make-constant-medium obj density albedo-texture)
  (make-obj 6 (m:make-lambertian albedo-texture) (list density) (list obj)))
