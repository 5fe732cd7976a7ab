;; A scene script in the reference's style (main.scm:31-89, 316-398): it only uses the reference's
;; constructor API, (use ...) lines and all, and ends with (srt:write-scene ...) instead of (trace-all ...).
(define-module demo
  (use srfi-27)
  (use vec :prefix v:)
  (use geometry :prefix g:)
  (use material :prefix m:)
  (use texture :prefix t:)
  (use bezier :prefix b:)
  (use camera :prefix cam:)
  (use srt-scene :prefix srt:)
  (export demo-scene))
(select-module demo)

(define *camera*
  (let ((lookfrom (v:vec3 13 2 3))
        (lookat (v:vec3 0 0 0)))
    (cam:make-camera lookfrom lookat (v:vec3 0 1 0) 20 (/ 3 2) 0.1 10 0 1)))

(define demo-scene
  (let* ((checker (t:checker-texture (t:constant-texture (v:vec3 0.2 0.3 0.1))
                                     (t:constant-texture (v:vec3 0.9 0.9 0.9))))
         (white (m:make-lambertian (t:constant-texture (v:vec3 0.73 0.73 0.73))))
         (marble (m:make-lambertian (t:marble-texture 0.25)))
         (noise (m:make-lambertian (t:noise-texture 4)))
         (light (m:make-diffuse-light (t:constant-texture (v:vec3 4 4 4))))
         (image (m:make-lambertian (t:image-texture (vector 255 0 0  0 255 0  0 0 255  10 20 30  40 50 60  70 80 90) 3 2)))
         (box (g:translate (g:rotate-y (g:make-box (v:vec3 0 0 0) (v:vec3 1 2 1) white) 15)
                           (v:vec3 -3 0 1)))
         (net (map (lambda (i)
                     (map (lambda (j) (v:vec3 i (* 0.25 i j) j)) '(0 1 2 3)))
                   '(0 1 2 3))))
    (g:make-scene
     (list (g:make-sphere (v:vec3 0 -1000 0) 1000 (m:make-lambertian checker))
           (g:make-sphere (v:vec3 0 1 0) 1 (m:make-dielectric 1.5))
           (g:make-sphere (v:vec3 0 1 0) -0.95 (m:make-dielectric 1.5))
           (g:make-sphere (v:vec3 4 1 0) 1 (m:make-metal (t:constant-texture (v:vec3 0.7 0.6 0.5)) 0.25))
           (g:make-moving-sphere (v:vec3 -4 1 0) (v:vec3 -4 1.5 0) 0 1 0.5 marble)
           (g:flip-normals (g:make-xz-rect -1 1 -1 1 5 light))
           (g:make-xy-rect 3 5 1 3 -2 noise)
           (g:flip-normals (g:make-yz-rect 0 2 0 2 6 image))
           box
           (g:make-constant-medium (g:translate (g:make-box (v:vec3 0 0 0) (v:vec3 1 1 1) white) (v:vec3 2 0 2))
                                   0.5 (t:constant-texture (v:vec3 1 1 1)))
           (g:make-bvh-node
            (list (b:make-bezier (v:vec3 -1 0 -1) (v:vec3 -0.8 1 1) (v:vec3 0.8 -1 1) (v:vec3 1 0 -1) 0.1 white)
                  (g:make-sphere (v:vec3 2 0.25 2) 0.25 (m:make-isotropic (t:constant-texture (v:vec3 0.5 0.5 0.5)))))
            0 1)
           (b:make-bezier-patch net white)
           (g:make-klein (v:vec3 250 200 280) white))
     *camera*
     srt:sky-color)))
