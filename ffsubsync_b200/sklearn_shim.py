"""Minimal fit/transform plumbing with the surface the reference's callers use.

Written from scratch for this package; it keeps the behaviour of the reference's shim
(ffsubsync/sklearn_shim.py:52-80 TransformerMixin, :89-335 Pipeline, :362-386 make_pipeline)
that the hot path relies on:

  * ``TransformerMixin.fit_transform(X, y=None, **fit_params)`` = ``fit(...).transform(X)``,
    passing ``y`` only when it is not None;
  * ``Pipeline(steps)``: ``fit`` / ``fit_transform`` push X through every step but the last with
    fit_transform, then fit the last one; ``transform`` is a *property* returning a callable;
    ``steps``, ``named_steps``, ``len()``, indexing by int / slice / name; ``None`` or
    ``"passthrough"`` steps are skipped; ``stepname__param`` routing of fit parameters;
  * ``make_pipeline(*steps)`` names steps after their lower-cased class, numbering duplicates.
"""
from typing import Any, Dict, Iterator, List, Optional, Tuple


class TransformerMixin:
    def fit_transform(self, X: Any, y: Optional[Any] = None, **fit_params: Any) -> Any:
        fitted = self.fit(X, **fit_params) if y is None else self.fit(X, y, **fit_params)
        return fitted.transform(X)


def _skipped(step) -> bool:
    return step is None or (isinstance(step, str) and step == "passthrough")


def _fit_transform_step(step, X, y, params):
    if hasattr(step, "fit_transform"):
        return step.fit_transform(X, y, **params)
    return step.fit(X, y, **params).transform(X)


class Pipeline:
    def __init__(self, steps: List[Tuple[str, Any]], verbose: bool = False) -> None:
        self.steps = steps
        self.verbose = verbose
        self._check_steps()

    # -- validation ---------------------------------------------------------------------------
    def _check_steps(self) -> None:
        *body, (_, last) = self.steps
        for _, step in body:
            if _skipped(step):
                continue
            can_fit = hasattr(step, "fit") or hasattr(step, "fit_transform")
            if not can_fit or not hasattr(step, "transform"):
                raise TypeError(
                    "All intermediate steps should be transformers and implement fit and "
                    "transform or be the string 'passthrough' '%s' (type %s) doesn't"
                    % (step, type(step)))
        if not _skipped(last) and not hasattr(last, "fit"):
            raise TypeError(
                "Last step of Pipeline should implement fit or be the string 'passthrough'. "
                "'%s' (type %s) doesn't" % (last, type(last)))

    # -- container protocol ---------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.steps)

    def __getitem__(self, ind):
        if isinstance(ind, slice):
            if ind.step not in (1, None):
                raise ValueError("Pipeline slicing only supports a step of 1")
            return self.__class__(self.steps[ind])
        if isinstance(ind, str):
            return self.named_steps[ind]
        return self.steps[ind][1]

    @property
    def named_steps(self) -> Dict[str, Any]:
        return dict(self.steps)

    @property
    def _final_estimator(self):
        last = self.steps[-1][1]
        return "passthrough" if last is None else last

    def _active(self, with_final: bool = True) -> Iterator[Tuple[int, str, Any]]:
        stop = len(self.steps) if with_final else len(self.steps) - 1
        for i in range(stop):
            name, step = self.steps[i]
            if not _skipped(step):
                yield i, name, step

    # -- estimator interface ----------------------------------------------------------------------
    def _route(self, fit_params: Dict[str, Any]) -> Dict[str, Dict[str, Any]]:
        routed: Dict[str, Dict[str, Any]] = {name: {} for name, step in self.steps if step is not None}
        for key, value in fit_params.items():
            if "__" not in key:
                raise ValueError(
                    "Pipeline.fit does not accept the {} parameter. You can pass parameters to "
                    "specific steps of your pipeline using the stepname__parameter format, e.g. "
                    "`Pipeline.fit(X, y, logisticregression__sample_weight=sample_weight)`."
                    .format(key))
            step, param = key.split("__", 1)
            routed[step][param] = value
        return routed

    def _fit_body(self, X, y, fit_params):
        self.steps = list(self.steps)
        self._check_steps()
        routed = self._route(fit_params)
        for i, name, step in self._active(with_final=False):
            X = _fit_transform_step(step, X, y, routed[name])
            self.steps[i] = (name, step)
        last_name = self.steps[-1][0]
        return X, ({} if _skipped(self.steps[-1][1]) else routed[last_name])

    def fit(self, X, y=None, **fit_params) -> "Pipeline":
        Xt, last_params = self._fit_body(X, y, fit_params)
        if not _skipped(self.steps[-1][1]):
            self.steps[-1][1].fit(Xt, y, **last_params)
        return self

    def fit_transform(self, X, y=None, **fit_params):
        Xt, last_params = self._fit_body(X, y, fit_params)
        last = self.steps[-1][1]
        if _skipped(last):
            return Xt
        return _fit_transform_step(last, Xt, y, last_params)

    @property
    def transform(self):
        last = self._final_estimator
        if last != "passthrough":
            last.transform  # AttributeError here if the final step cannot transform
        return self._transform

    def _transform(self, X):
        for _, _, step in self._active():
            X = step.transform(X)
        return X


def make_pipeline(*steps, **kwargs) -> Pipeline:
    verbose = kwargs.pop("verbose", False)
    if kwargs:
        raise TypeError('Unknown keyword arguments: "{}"'.format(list(kwargs.keys())[0]))
    names = [s if isinstance(s, str) else type(s).__name__.lower() for s in steps]
    totals: Dict[str, int] = {}
    for n in names:
        totals[n] = totals.get(n, 0) + 1
    seen: Dict[str, int] = {}
    labelled = []
    for n, s in zip(names, steps):
        if totals[n] > 1:
            seen[n] = seen.get(n, 0) + 1
            labelled.append(("%s-%d" % (n, seen[n]), s))
        else:
            labelled.append((n, s))
    return Pipeline(labelled, verbose=verbose)
