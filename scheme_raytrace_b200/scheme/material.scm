;; material.scm -- material constructors that record table rows: #(material kind texture param)
(define-module material
  (export make-lambertian make-metal make-dielectric make-diffuse-light make-isotropic
          material-kind material-texture material-param))
(select-module material)

(define (mk kind tex param) (vector 'material kind tex param))
(define (material-kind m) (vector-ref m 1))
(define (material-texture m) (vector-ref m 2))
(define (material-param m) (vector-ref m 3))
(define (make-lambertian albedo) (mk 0 albedo 0))
(define (make-metal albedo fuzz) (mk 1 albedo fuzz))
(define (make-dielectric ref-idx) (mk 2 #f ref-idx))
(define (make-diffuse-light emit) (mk 3 emit 0))
(define (make-isotropic albedo) (mk 4 albedo 0))
