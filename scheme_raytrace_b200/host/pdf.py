"""pdf.scm — pdf CONSTRUCTORS.  Upstream these are closure vectors `#(value-fn generate-fn)`
evaluated inside the (never wired-up) Rest-of-Life estimator; here they only select the estimator
of the render call: `make_mixture_pdf(make_hitable_pdf(light, None), make_cosine_pdf(None))`
== estimator SRT_EST_MIXTURE with `lights=[prim id of light]`."""
from dataclasses import dataclass
from typing import Any

EST_REFERENCE, EST_MIXTURE = 0, 1


@dataclass(frozen=True)
class Pdf:
    kind: str
    a: Any = None
    b: Any = None


def make_cosine_pdf(w=None):                 # pdf.scm:18-26
    return Pdf("cosine", w)


def make_hitable_pdf(obj, origin=None):      # pdf.scm:28-32 (g:pdf-value / g:random are undefined upstream)
    return Pdf("hitable", obj, origin)


def make_mixture_pdf(p0, p1):                # pdf.scm:34-41
    return Pdf("mixture", p0, p1)


def estimator_of(pdf, flat):
    """(estimator, light primitive ids) for a pdf expression over a flattened scene."""
    if pdf is None or pdf.kind == "cosine":
        return EST_REFERENCE, []
    if pdf.kind == "mixture":
        lights = []
        for q in (pdf.a, pdf.b):
            if q.kind == "hitable":
                objs = q.a if isinstance(q.a, (list, tuple)) else [q.a]
                for o in objs:
                    leaf = o.children[0] if o.kind == 16 else o          # look through flip-normals
                    lights.append(next(i for i, l in enumerate(flat.leaves) if l is leaf))
        return EST_MIXTURE, lights
    raise ValueError("unsupported pdf expression")
