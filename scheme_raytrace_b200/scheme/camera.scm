;; camera.scm -- make-camera: the 10-slot vector (llc horiz vert origin w u v lens-radius time0 time1)
(define-module camera
  (use vec :prefix v:)
  (export make-camera))
(select-module camera)

(define (make-camera lookfrom lookat vup vfov aspect aperture focus-dist time0 time1)
  (let* ((half-h (tan (/ (* vfov (/ 3.141592653589793 180)) 2)))
         (half-w (* aspect half-h))
         (w (v:unit (v:diff lookfrom lookat)))
         (u (v:unit (v:cross vup w)))
         (v (v:cross w u))
         (llc (v:diff lookfrom
                      (v:scale u (* half-w focus-dist))
                      (v:scale v (* half-h focus-dist))
                      (v:scale w focus-dist))))
    (vector llc
            (v:scale u (* 2 half-w focus-dist))
            (v:scale v (* 2 half-h focus-dist))
            lookfrom w u v (/ aperture 2) time0 time1)))
