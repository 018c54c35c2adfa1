"""Drop-in switch for code that imports the reference package itself.

    import grates, grates_b200
    grates_b200.install()          # rebinds the hot methods of grates' own classes
    ...                            # user code keeps calling grates as before
    grates_b200.uninstall()

The reference has no plugin or operator registry: its boundary for this path is the
method signatures.  ``install`` replaces exactly those methods with wrappers that keep argument
meaning, return type (the reference's own grid / coefficient classes), side effects (``self.values``
after a covariance propagation) and exceptions, and route the arithmetic through the C ABI of
``include/grates_b200.h``.  Nothing else of grates is touched; without a CUDA device the wrappers
raise (there is no CPU fallback).

Rebound (reference file:line):
    PotentialCoefficients.to_grid                 gravityfield.py:331-390
    RegularGrid.to_potential_coefficients         grid.py:752-790
    RegularGrid.covariance_propagation            grid.py:792-839
    IrregularGrid.covariance_propagation          grid.py:1071-1120
    OrderWiseFilter.filter / Gaussian.filter / Butterworth.filter      filter.py:43-70, 108-118, 153-191
    GeneralMatrix.filter / VDK.filter             filter.py:456-479, 548-572
    gravityfield.gridded_rms                      gravityfield.py:1143-1172
    RadialBasisFunctions.to_potential_coefficients      gravityfield.py:692-727
"""
import numpy as np

from . import filter as _filter, gravityfield as _gf, grid as _grid

_ORIGINALS = []       # (owner, attribute name, original)


def _mirror_grid(grid):
    """grates_b200 grid with the geometry of a reference grid (regular if it has parallels / meridians)."""
    a, f = grid.semimajor_axis, grid.flattening
    if hasattr(grid, "parallels") and hasattr(grid, "meridians"):
        area = grid.area
        if area is not None:
            area = np.asarray(area, dtype=float).reshape(grid.parallels.size, grid.meridians.size)
        return _grid.RegularGrid(grid.meridians, grid.parallels, area, a, f)
    return _grid.IrregularGrid(grid.longitude, grid.latitude, grid.area, a, f)


def _mirror_coefficients(pc):
    out = _gf.PotentialCoefficients(pc.GM, pc.R)
    out.anm = np.ascontiguousarray(pc.anm, dtype=float)
    out.epoch = getattr(pc, "epoch", None)
    return out


def _rebind(owner, name, new):
    _ORIGINALS.append((owner, name, getattr(owner, name)))
    setattr(owner, name, new)


def install(reference=None):
    """Rebind the hot-path methods of the reference package (default: ``import grates``)."""
    if reference is None:
        import grates as reference
    if _ORIGINALS:
        return reference
    ref = reference

    def to_grid(self, grid=None, kernel='ewh'):
        grid = ref.grid.GeographicGrid() if grid is None else grid
        values = _mirror_coefficients(self).to_grid(_mirror_grid(grid), kernel).values
        out = grid.copy()                  # keeps the grid's own epoch, as gravityfield.py:367 does
        out.values = np.ascontiguousarray(values).reshape(-1)
        return out

    def to_potential_coefficients(self, min_degree, max_degree, kernel='potential', GM=3.9860044150e+14,
                                  R=6.3781363000e+06):
        if self.values is None:
            raise ValueError("grid has no values")
        mine = _mirror_grid(self)
        mine.values = np.ascontiguousarray(self.values, dtype=float).reshape(-1)
        res = mine.to_potential_coefficients(min_degree, max_degree, kernel, GM, R)
        out = ref.gravityfield.PotentialCoefficients(GM, R)     # epoch stays None, as in grid.py:752-790
        out.anm = res.anm
        return out

    def covariance_propagation(self, covariance_matrix, min_degree, max_degree, kernel='potential',
                               GM=3.9860044150e+14, R=6.3781363000e+06):
        std = _mirror_grid(self).covariance_propagation(covariance_matrix, min_degree, max_degree, kernel, GM, R)
        self.values = np.ascontiguousarray(std).reshape(-1)
        return self.values.copy()

    def _filter_with(make):
        def filter(self, gravityfield):
            if not isinstance(gravityfield, ref.gravityfield.PotentialCoefficients):
                raise TypeError("Filter operation only implemented for instances of 'PotentialCoefficients'")
            res = make(self).filter(_mirror_coefficients(gravityfield))
            out = gravityfield.copy()
            out.anm = res.anm
            return out
        return filter

    def dense_filter(self, gravityfield):
        # GeneralMatrix / VDK keep their matrix under name-mangled attributes; the tiled copy is cached on the object
        mine = self.__dict__.get("_gb_mirror")
        if mine is None:
            mine = _filter.GeneralMatrix(self._GeneralMatrix__W, self._GeneralMatrix__nmin, self._GeneralMatrix__nmax)
            self.__dict__["_gb_mirror"] = mine
        res = mine.filter(_mirror_coefficients(gravityfield))
        out = gravityfield.copy()
        out.anm = res.anm
        return out

    def rbf_to_potential_coefficients(self, blocking_factor=256):
        mine = self.__dict__.get("_gb_mirror")
        if mine is None:
            mine = _gf.RadialBasisFunctions(_mirror_grid(self.point_distribution), self._RadialBasisFunctions__K,
                                            self._RadialBasisFunctions__min_degree,
                                            self._RadialBasisFunctions__max_degree, self.GM, self.R)
            self.__dict__["_gb_mirror"] = mine
        mine.GM, mine.R, mine.epoch = self.GM, self.R, self.epoch
        mine.values = np.ascontiguousarray(self.values, dtype=float)
        res = mine.to_potential_coefficients()
        out = ref.gravityfield.PotentialCoefficients(self.GM, self.R)
        out.anm = res.anm
        out.epoch = self.epoch
        return out

    def gridded_rms(temporal_gravityfield, epochs, kernel='ewh', base_grid=None):
        base_grid = ref.grid.GeographicGrid() if base_grid is None else base_grid

        class _Series:            # evaluate_at through the reference object, synthesis batched on the GPU
            def evaluate_at(self, t):
                return _mirror_coefficients(temporal_gravityfield.evaluate_at(t))
        values = _gf.gridded_rms(_Series(), epochs, kernel, _mirror_grid(base_grid)).values
        out = base_grid.copy()
        out.values = np.ascontiguousarray(values).reshape(-1)
        return out

    try:
        _rebind(ref.gravityfield.PotentialCoefficients, "to_grid", to_grid)
        _rebind(ref.grid.RegularGrid, "to_potential_coefficients", to_potential_coefficients)
        _rebind(ref.grid.RegularGrid, "covariance_propagation", covariance_propagation)
        _rebind(ref.grid.IrregularGrid, "covariance_propagation", covariance_propagation)
        _rebind(ref.filter.OrderWiseFilter, "filter",
                _filter_with(lambda f: _filter.OrderWiseFilter(f._OrderWiseFilter__array)))
        _rebind(ref.filter.Gaussian, "filter", _filter_with(lambda f: _filter.Gaussian(f.radius)))
        _rebind(ref.filter.Butterworth, "filter", _filter_with(lambda f: _filter.Butterworth(f.order, f.cutoff_degree)))
        if hasattr(ref.filter, "GeneralMatrix"):
            _rebind(ref.filter.GeneralMatrix, "filter", dense_filter)
            _rebind(ref.filter.VDK, "filter", dense_filter)      # the reference's own VDK.filter raises AttributeError
        _rebind(ref.gravityfield, "gridded_rms", gridded_rms)
        if hasattr(ref.gravityfield, "RadialBasisFunctions"):
            _rebind(ref.gravityfield.RadialBasisFunctions, "to_potential_coefficients", rbf_to_potential_coefficients)
    except Exception:
        uninstall()
        raise
    return reference


def uninstall():
    """Restore the reference's own methods."""
    for owner, name, original in reversed(_ORIGINALS):
        setattr(owner, name, original)
    _ORIGINALS.clear()


def installed():
    return sorted("{0}.{1}".format(getattr(owner, "__name__", type(owner).__name__), name) for owner, name, _ in _ORIGINALS)
