"""CPU: self-checks of oracle/minischeme.py, the interpreter that executes the reference for the golden vectors.

The interpreter is NOT Gauche; these cases pin the semantics the reference's sources rely on against values that
are fixed by R7RS / the Gauche manual (numeric tower, multiple values, hygiene of syntax-rules, the Gauche forms
`receive`, `let-values`, `dotimes`, `inc!`, `push!`, `let-optionals*`, uniform vectors, `array-mul`)."""
from fractions import Fraction
import math
import pytest
from oracle.minischeme import Interp, read_all, F64, Values


def run(src, random_real=None):
    it = Interp([], random_real or (lambda: 0.5))
    out = None
    for form in read_all(src):
        out = it.eval(form, it.user)
    return out


def lst(x):
    return list(x)


@pytest.mark.parametrize("src, want", [
    ("(+ 1 2 3)", 6), ("(/ 1 3)", Fraction(1, 3)), ("(/ 6 3)", 2), ("(/ 1 3.0)", 1 / 3.0), ("(* 1/2 4)", 2), ("(- 5)", -5),
    ("(/ 1 0.0)", math.inf), ("(/ -1 0.0)", -math.inf), ("(exact->inexact 1/3)", 1 / 3), ("(sqrt 16)", 4), ("(sqrt 2)", math.sqrt(2)),
    ("(sqrt 1/4)", Fraction(1, 2)), ("(floor->exact 2.7)", 2), ("(floor->exact -2.5)", -3), ("(max 1 2.0)", 2.0), ("(min 1 2)", 1),
    ("(abs -7/2)", Fraction(7, 2)), ("(expt 2 10)", 1024), ("(quotient 7 2)", 3), ("(remainder -7 2)", -1), ("(modulo -7 2)", 1),
    ("(atan 1 1)", math.atan2(1, 1)), ("(* 2 pi)", 2 * math.pi), ("(< 1 2 3)", True), ("(< 1 3 2)", False), ("(= 1 1.0)", True),
    ("(let ((x 2) (y 3)) (* x y))", 6), ("(let* ((x 2) (y (* x 3))) y)", 6), ("(letrec ((f (lambda (n) (if (= n 0) 1 (* n (f (- n 1))))))) (f 5))", 120),
    ("(let loop ((i 0) (acc 0)) (if (< i 5) (loop (+ i 1) (+ acc i)) acc))", 10),
    ("(do ((i 0 (+ i 1)) (s 0 (+ s i))) ((= i 4) s))", 6),
    ("(let ((s 0)) (dotimes (i 4) (set! s (+ s i))) s)", 6),
    ("(let ((x 1)) (inc! x) (inc! x 3) x)", 5), ("(let ((x 5)) (dec! x) x)", 4),
    ("(let ((l '())) (push! l 1) (push! l 2) (car l))", 2),
    ("(call-with-values (lambda () (values 1 2)) +)", 3), ("(receive (a b) (values 1 2) (- a b))", -1),
    ("(receive (a . rest) (values 1 2 3) (length rest))", 2), ("(let-values (((a b) (values 1 2)) ((c) (values 3))) (+ a b c))", 6),
    ("(let*-values (((a b) (values 1 2)) ((c) (values (+ a b)))) c)", 3),
    ("(+ (values 1 2) 10)", 11),                                  # Gauche: extra values are dropped in a one-value context
    ("(cond ((> 1 2) 'a) ((> 2 1) 'b) (else 'c))", "b"),
    ("(case 3 ((1 2) 'low) ((3 4) 'mid) (else 'high))", "mid"), ("(and 1 2 #f 3)", False), ("(or #f #f 3)", 3), ("(if '() 1 2)", 1),
    ("(when (> 2 1) 1 2)", 2), ("(unless (< 2 1) 1 2)", 2),
    ("(vector-ref (vector 1 2 3) 1)", 2), ("(let ((v (make-vector 3 0))) (vector-set! v 1 9) (vector-ref v 1))", 9),
    ("(apply + 1 2 '(3 4))", 10), ("(length (map (lambda (x y) (+ x y)) '(1 2 3) '(4 5 6)))", 3), ("(car (reverse '(1 2 3)))", 3),
    ("(f64vector-ref (f64vector-add (f64vector 1 2 3) (f64vector 10 20 30)) 2)", 33.0), ("(f64vector-dot (f64vector 1 2 3) (f64vector 4 5 6))", 32.0),
    ("(f64vector-ref (f64vector-mul (f64vector 1 2 3) 2) 1)", 4.0), ("(f64vector-ref (f64vector-div (f64vector 1 2 3) (f64vector 2 2 2)) 2)", 1.5),
    ("(let-optionals* '(7) ((a 1) (b 2)) (+ a b))", 9), ("(let1 x 4 (* x x))", 16),
    ("(string-append \"a\" \"b\")", "ab"), ("(format \"~D ~D\\n\" 3 4)", "3 4\n"), ("(string->number \"2.5\")", 2.5), ("(length (string-split \"1,2,3\" \",\"))", 3),
])
def test_expression(src, want):
    got = run(src)
    if isinstance(want, float):
        assert isinstance(got, float) and (got == want or abs(got - want) < 1e-15)
    else:
        assert got == want and (type(got) is type(want) or isinstance(want, (bool, str, type(None))))


def test_syntax_rules_is_hygienic_and_duplicates_operands():
    # (a) a template identifier refers to the binding at the macro's DEFINITION, even if the use site shadows it
    assert run("(define-syntax my-or (syntax-rules () ((_ a b) (let ((t a)) (if t t b))))) (let ((t 5)) (my-or #f t))") == 5
    assert run("(define (helper x) (* 10 x)) (define-syntax call-helper (syntax-rules () ((_ x) (helper x)))) (let ((helper (lambda (x) 0))) (call-helper 4))") == 40
    # (b) an operand mentioned three times in the template is EVALUATED three times - what makes Q15 (onb.scm:27-36)
    draws = iter([0.1, 0.2, 0.3, 0.4])
    src = "(define-syntax three (syntax-rules () ((_ a) (list a a a)))) (three (random-real))"
    assert lst(run(src, lambda: next(draws))) == [0.1, 0.2, 0.3]
    # (c) several rules, chosen by operand count (the two forms of `local`)
    assert run("(define-syntax f (syntax-rules () ((_ a b c) (+ a b c)) ((_ a) (* a a)))) (+ (f 1 2 3) (f 4))") == 22


def test_argument_evaluation_is_left_to_right():
    draws = iter([1.0, 2.0, 3.0])
    assert lst(run("(list (random-real) (random-real) (random-real))", lambda: next(draws))) == [1.0, 2.0, 3.0]


def test_modules_prefix_and_export(tmp_path):
    (tmp_path / "m1.scm").write_text("(define-module m1 (export f g)) (select-module m1) (define (f x) (+ x 1)) (define (g x) (* (f x) 2)) (define hidden 3)")
    (tmp_path / "m2.scm").write_text("(define-module m2 (use m1 :prefix a:) (export-all)) (select-module m2) (define (h x) (a:g (a:f x)))")
    it = Interp([str(tmp_path)])
    it.require("m2")
    assert it.call("m2", "h", 1) == 6
    with pytest.raises(Exception):
        it.eval(read_all("(a:hidden)")[0], it.modules["m2"])


def test_array_mul_and_closures_in_vectors():
    # gauche.array: (array (shape 0 r 0 c) ...) and array-mul, as bezier.scm uses them for its ray-space transform
    src = "(let ((a (array (shape 0 2 0 2) 1 2 3 4)) (b (array (shape 0 2 0 1) 5 6))) (let ((c (array-mul a b))) (list (array-ref c 0 0) (array-ref c 1 0))))"
    assert lst(run(src)) == [17, 39]
    # the reference's object protocol: closures stored in vectors, called through vector-ref
    assert run("(define obj (vector (lambda (x) (* x 3)) 'tag)) ((vector-ref obj 0) 7)") == 21
