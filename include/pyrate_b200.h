/*
 * pyrate_b200 -- C ABI of the B200-native sequential trace engine.
 *
 * This is the drop-in boundary for ONE path of mess42/pyrate (pyrateoptics
 * 0.4.0): OpticalSystem.seqtrace (raytracer/optical_system.py:73-94) ->
 * OpticalElement.seqtrace (raytracer/optical_element.py:324-379) and the
 * per-surface step below it.  The reference has no FFI on this path (it is pure
 * NumPy); its one FFI precedent is the ctypes binding of a user DLL in
 * raytracer/surface_shape_zmxdll.py:293, and INTEGRATION.md shows the ctypes
 * stub a pyrate maintainer would add inside OpticalSystem.seqtrace to call
 * pyr_trace().  Signatures are plain C: pointers, sizes, PODs.  No torch types.
 *
 * Conventions
 *   - all ray arrays are component-major "(3, n)" like the reference's
 *     RayBundle rows (raytracer/ray.py:40-66): element (c, i) lives at
 *     base[c * ld + i] with leading dimension ld >= n.  Inputs with ld even and
 *     16-byte aligned rows are staged by the TMA (reads may touch one pad element
 *     past n); outputs with ld a multiple of 16 and 16-byte aligned rows leave as
 *     TMA bulk stores (writes may touch the row pad up to the next multiple of 16
 *     elements).  Anything else still works through plain loads / stores.
 *   - complex arrays are interleaved (re, im) doubles = numpy/torch complex128;
 *     element (c, i) at base[2 * (c * ld + i)].
 *   - every device pointer is caller-owned; calls are asynchronous on `stream`.
 *     The only device memory the library itself touches is the 8-byte tile
 *     cursor of a large launch (>= 8 tiles per resident CTA of the conic-only
 *     kernels): borrowed from a private stream-ordered pool
 *     (cudaMallocFromPoolAsync), zeroed, used and released in stream order
 *     -- calls stay re-entrant per stream and per host thread; if the pool is
 *     unavailable the launch uses the static tile schedule instead.
 *   - return value: 0 = ok, < 0 = PYR_E_* (see pyr_strerror), > 0 = cudaError_t.
 *   - per-ray failures are never errors: they are reported through the flag
 *     byte (and NaN coordinates), like the reference's `valid` rows.
 */
#ifndef PYRATE_B200_H
#define PYRATE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYR_ABI_VERSION 4

#define PYR_MAX_COEFF 80        /* asphere coefficients / XY-polynomial terms (a
                                   Zernike series up to Fringe term 36 expands into
                                   66 monomials)                                   */
#define PYR_MAX_GRIN_PARAMS 8
#define PYR_MAX_WAVES 4         /* wavelength segments of one pyr_trace call        */
#define PYR_MAX_TERMS 4         /* sub-shapes of a PYR_SHAPE_COMBINATION           */

/* error codes */
#define PYR_OK 0
#define PYR_E_BADARG (-1)
#define PYR_E_UNSUPPORTED (-2)
#define PYR_E_TOOLARGE (-3)
#define PYR_E_NODEVICE (-4)

/* Shape.intersect variants (raytracer/surface_shape.py) */
enum PyrShapeKind {
    PYR_SHAPE_CONIC = 0,        /* Conic.intersect :289-325, closed form          */
    PYR_SHAPE_ASPHERE = 1,      /* ExplicitShape.intersect :448-465 + Asphere.F   */
    PYR_SHAPE_XYPOLY = 2,       /* ExplicitShape.intersect + XYPolynomials.F :785 */
    PYR_SHAPE_BICONIC = 3,      /* ExplicitShape.intersect + Biconic.F :618-629:
                                   curv/cc = x section, curv2/cc2 = y section,
                                   coeff[i] = A_(2i+2), coeff[16 + i] = B_(2i+2),
                                   n_coeff <= 16 pairs                             */
    PYR_SHAPE_GRIDSAG = 4,      /* ExplicitShape.intersect + GridSag.F :866-880: the
                                   bicubic tensor-product B-spline scipy's
                                   RectBivariateSpline (FITPACK) fits through the
                                   sag grid, evaluated like its ev(): arguments
                                   clamped to the grid, de Boor basis             */
    PYR_SHAPE_COMBINATION = 5,  /* ExplicitShape.intersect + LinearCombination.F
                                   :714-731: z = sum_i w_i (F_i(x - dx_i, y - dy_i)
                                   + dz_i) over PyrStep.terms                      */
    PYR_SHAPE_CYLINDER = 6      /* Cylinder :328-388: conic section in y, extruded
                                   along x: c (y^2 + (1 + cc) z^2) - 2 z = 0, closed
                                   form.  The reference's intersect is dead code
                                   (it reads a non-existent attribute, :380) and its
                                   H omits d_x; this is the corrected quadratic:
                                   H = -c (d_y^2 + (1 + cc) d_z^2)                 */
};

/* One sub-shape of a PYR_SHAPE_COMBINATION.  Its frame differs from the combination's
 * by a translation (dx, dy, dz) only (the Zemax importer's decentred Zernike term,
 * io/zmx.py:755-775).  Coefficients are the slice [coeff_off, coeff_off + coeff_len)
 * of the step's coeff / xpow / ypow arrays in the layout of `kind`; a GRIDSAG term
 * uses the step's grid_* arrays (at most one per step).  A Conic sub-shape is passed
 * as ASPHERE with n_coeff = 0.                                                      */
typedef struct PyrShapeTerm {
    int32_t kind;               /* ASPHERE, XYPOLY, BICONIC or GRIDSAG             */
    int32_t n_coeff;
    int32_t coeff_off;
    int32_t coeff_len;
    double weight;
    double dx, dy, dz;
    double curv, cc, curv2, cc2;
    double normradius;
} PyrShapeTerm;

/* raytracer/aperture.py:71-140 */
enum PyrApertureKind {
    PYR_AP_BASE = 0,            /* passes everything                               */
    PYR_AP_CIRCULAR = 1,        /* p[0] = minradius, p[1] = maxradius              */
    PYR_AP_RECTANGULAR = 2      /* p[0] = width, p[1] = height                     */
};

/* what the material on the far side does with the ray (optical_element.py:338-368) */
enum PyrInteraction {
    PYR_REFRACT = 0,            /* Material.refract                                */
    PYR_REFLECT = 1             /* Material.reflect (is_mirror)                    */
};

/* medium model (raytracer/material/*.py) */
enum PyrMediumKind {
    PYR_MEDIUM_ISO_CONST = 0,   /* IsotropicMaterial with position-independent n   */
    PYR_MEDIUM_ISO_GRIN = 1,    /* IsotropicGrinMaterial (material_grin.py)        */
    PYR_MEDIUM_ANISO = 2        /* AnisotropicMaterial (material_anisotropic.py)   */
};

/* closed catalogue of GRIN index profiles, evaluated in the material frame */
enum PyrGrinProfile {
    PYR_GRIN_GAUSSIAN_XY = 0,   /* n = p0 + p1 exp(-p2 x^2 - p3 y^2)               */
    PYR_GRIN_POLY_RZ = 1,       /* n = p0 + p1 r^2 + p2 r^4 + p3 r^6
                                       + p4 z + p5 z^2 + p6 z^3,  r^2 = x^2+y^2    */
    PYR_GRIN_USER = 100         /* index function given as CUDA source by the user
                                   (core/functionobject.py:99-119 of the reference
                                   takes Python source): such a medium is integrated
                                   by a kernel compiled at run time (NVRTC,
                                   pyrate_b200/grin_jit.py); this library refuses it
                                   as `before` (PYR_E_UNSUPPORTED) and, as `after`,
                                   needs the per-ray index in PyrStep.after_n_rays    */
};

enum PyrGrinBoundary {
    PYR_BND_NONE = 0,
    PYR_BND_CYLINDER = 1,       /* x^2 + y^2 < b0^2                                */
    PYR_BND_BOX = 2,            /* |x| < b0 and |y| < b1                           */
    PYR_BND_SPHERE = 3          /* x^2 + y^2 + z^2 < b0^2                          */
};

/* which half of the step runs: the fused path always uses PYR_STEP_FULL; the two
 * partial modes back the reference's stand-alone plugin calls
 * Material.propagate / Surface.intersect and Material.refract / reflect, in the
 * real-valued and in the complex-valued (PYR_F_COMPLEX: crystals, complex k) kernels;
 * a splitting PYR_STEP_DEFLECT_ONLY step writes both modes (width 2n, ld_out2)       */
enum PyrStepMode {
    PYR_STEP_FULL = 0,
    PYR_STEP_PROPAGATE_ONLY = 1, /* intersect + aperture; k, E unchanged            */
    PYR_STEP_DEFLECT_ONLY = 2    /* x is already the hit point; refract / reflect   */
};

/* how the ray direction d is obtained from (k, E) before the intersect
 * (RayBundle.returnKtoD, raytracer/ray.py:136-152) */
enum PyrDirMode {
    PYR_DIR_POYNTING = 0,       /* d = S/|S|, S = Re(|E|^2 k - (E.k) conj(E))      */
    PYR_DIR_K = 1               /* d = Re(k)/|Re(k)|: exact whenever E.k = 0, which
                                   every isotropic refraction guarantees          */
};

/* A frame maps local -> global as  x_g = R x_l + o  (R row-major, orthonormal),
 * global -> local as  x_l = R^T (x_g - o)   (localcoordinates.py:383-413).      */
typedef struct PyrFrame {
    double r[9];
    double o[3];
} PyrFrame;

typedef struct PyrMedium {
    int32_t kind;               /* PyrMediumKind                                   */
    int32_t grin_profile;       /* PyrGrinProfile                                  */
    int32_t grin_boundary;      /* PyrGrinBoundary                                 */
    int32_t grin_max_steps;     /* safety cap per ray (0 = default 1000000)        */
    double n;                   /* ISO_CONST: refractive index at the bundle wave  */
    double eps[18];             /* ANISO: complex 3x3 (row-major, re/im) in the
                                   MATERIAL frame (get_epsilon_tensor)             */
    double grin_p[PYR_MAX_GRIN_PARAMS];
    double grin_b[4];
    double grin_ds;             /* annotations["ds"] (material_grin.py:218)        */
    double grin_energy_tol;     /* annotations["energyviolation"], tested PER RAY  */
    PyrFrame frame;             /* material frame (material.lc)                    */
} PyrMedium;

/* One entry of the element sequence = propagate to the surface, intersect,
 * aperture, then refract/reflect into `after` (optical_element.py:336-375).     */
typedef struct PyrStep {
    int32_t shape_kind;         /* PyrShapeKind                                    */
    int32_t aperture_kind;      /* PyrApertureKind                                 */
    int32_t interaction;        /* PyrInteraction                                  */
    int32_t dir_mode;           /* PyrDirMode for the segment ending here          */
    int32_t n_coeff;            /* used entries of coeff[]                         */
    int32_t newton_maxit;       /* iteration cap of the explicit-shape solve       */
    int32_t split;              /* ANISO deflection only: 1 = this step doubles the
                                   rays (splitup=False, material_anisotropic.py
                                   :87-100); only allowed on the LAST step of a
                                   pyr_trace call                                  */
    int32_t mode;               /* PyrStepMode                                     */
    double k_norm_hint;         /* > 0: |Re k| on entry is known to be this value
                                   (index of `before` after an isotropic
                                   deflection); saves the normalisation in
                                   PYR_DIR_K mode.  0 = unknown                    */
    double curv, cc;            /* conic / asphere base conic; biconic x section   */
    double curv2, cc2;          /* biconic y section                               */
    double normradius;          /* XY polynomial                                   */
    double newton_tol;          /* |dt| <= tol (1 + |t|) ends the iteration        */
    double coeff[PYR_MAX_COEFF];        /* asphere: A2, A4, ...; XY: c_mn          */
    int8_t xpow[PYR_MAX_COEFF];         /* XY polynomial exponents                 */
    int8_t ypow[PYR_MAX_COEFF];
    double aperture_p[4];
    PyrFrame shape_frame;       /* shape.lc                                        */
    PyrFrame aperture_frame;    /* aperture.lc                                     */
    PyrMedium before;           /* medium the ray propagates in up to the surface  */
    PyrMedium after;            /* medium that deflects the ray at the surface     */

    /* where to record this step (device pointers; NULL = do not record).
     * n = rays of the call, n_out = n, or 2n on a split step.                    */
    double *out_x;              /* (3, n)      hit point, global frame             */
    double *out_k;              /* (3, n_out)  wave vector after deflection, global;
                                   complex if PYR_F_COMPLEX                        */
    double *out_e;              /* (3, n_out)  E field after deflection, global    */
    uint8_t *out_flags;         /* (n)         PYR_RAY_* bits                      */
    int64_t ld_out;             /* leading dimension of out_x / out_k / out_e      */
    /* optional GRIN integrator history of the segment ending at this surface (the rows
     * the reference appends per integrator step, material_grin.py:198-205); all NULL =
     * off.  Row r of ray i: hist_x[(r*3 + c)*ld_out + i] (frozen position, global),
     * hist_k likewise (k = p/n there), hist_valid[r*ld_out + i]; hist_count[i] = rows
     * the ray produced (<= hist_rows; later rows of the lock-step reference repeat the
     * last one).                                                                     */
    double *grin_hist_x;
    double *grin_hist_k;
    uint8_t *grin_hist_valid;
    int32_t *grin_hist_count;
    int64_t grin_hist_rows;
    /* PYR_SHAPE_GRIDSAG: FITPACK representation (device pointers, caller-owned):
     * knots grid_tx[grid_nx], grid_ty[grid_ny] (cubic: 4-fold end knots), coefficients
     * grid_c[(grid_nx - 4) * (grid_ny - 4)], x index major                            */
    const double *grid_tx;
    const double *grid_ty;
    const double *grid_c;
    int32_t grid_nx, grid_ny;
    /* PYR_SHAPE_COMBINATION */
    int32_t n_terms;
    int32_t reserved0;
    PyrShapeTerm terms[PYR_MAX_TERMS];
    int64_t ld_out2;            /* split step only: leading dimension of out_k /
                                   out_e (width 2n: mode a of ray i in column i,
                                   mode b in column n + i, the reference's hstack
                                   order, material_anisotropic.py:89-100)          */
    /* wavelength batch (PyrRaysIn.n_waves > 1): index of the ISO_CONST media `before` /
     * `after` for the rays of wavelength segment w (entry 0 repeats before.n / after.n;
     * dispersion is the only thing that differs between the bundles of a batch: F, d, C
     * bundles of one system, demos/demo_doublegauss.py:189-213, in ONE launch)         */
    double before_n_w[PYR_MAX_WAVES];
    double after_n_w[PYR_MAX_WAVES];
    /* optional (DEVICE, n doubles): refractive index of the deflecting medium AT THE HIT
     * POINT of every ray, evaluated by the caller -- position dependent media whose index
     * function this library does not know (PYR_GRIN_USER).  NULL = after.n / the catalogue
     * profile of after.                                                                  */
    const double *after_n_rays;
} PyrStep;

/* per-ray flag bits written to out_flags */
#define PYR_RAY_HIT 1u          /* still valid after propagate+intersect+aperture
                                   (RayBundle.valid[-1] after Surface.intersect)   */
#define PYR_RAY_ALIVE 2u        /* survives the deflection: contained in the next
                                   RayBundle (material_isotropic.py:183-199)       */

/* ---------------------------------------------------------------------------
 * Bundle generation on the device (the callers right above seqtrace:
 * OpticalSystemAnalysis.collimated_bundle / divergent_bundle,
 * raytracer/analysis/optical_system_analysis.py:83-165, over the rasters of
 * sampling2d/raster.py:36-166).  A bundle is fully described by (raster, radius, start
 * point, direction, background index): ray i of the call is raster point `first + i`,
 * computed in registers -- no start points, wave vectors or fields are read from
 * memory.  The reference runs a generalised eigen-solve per ray here even in vacuum
 * (material/material.py:476-497); in a homogeneous isotropic background the result is
 * k = n d and any unit E perpendicular to it.
 * ------------------------------------------------------------------------- */
enum PyrRasterKind {
    PYR_RASTER_HEXAPOLAR = 0,   /* ring j = 1..R carries 6 j points at radius j / R, angle
                                   2 pi i / (6 j); point 0 is the centre; 1 + 3 R (R + 1)
                                   points (BASELINE.json's "hexapolar-sampled bundles");
                                   param = R                                          */
    PYR_RASTER_RECT = 1,        /* RectGrid raster.py:36-60: square lattice x1d x x1d,
                                   x1d = linspace(lin_start, lin_stop, param), clipped to
                                   the unit disk; row-major (y index slow)             */
    PYR_RASTER_HEX = 2,         /* HexGrid :62-91: lattice x1d x sqrt(3) x1d followed by
                                   the same lattice shifted by (aux[0], aux[1]) = half a
                                   cell, both clipped to the unit disk; param = nx     */
    PYR_RASTER_CIRCULAR = 3     /* CircularGrid :150-166: param radii x param angles,
                                   angle index slow; PYR_GEN_SQRT_R: r -> sqrt(r)      */
};

enum PyrBundleKind {
    PYR_BUNDLE_COLLIMATED = 0,  /* x = radius * p + start, d = dir        (:83-125)    */
    PYR_BUNDLE_DIVERGENT = 1    /* x = start, d from the angles angley + radius p_x,
                                   anglex + radius p_y (:127-165); dir[0] = angley,
                                   dir[1] = anglex                                     */
};

#define PYR_GEN_E_PERP 1u       /* E = unit vector perpendicular to d built from the
                                   coordinate axis least aligned with d (else: e[])    */
#define PYR_GEN_SQRT_R 2u       /* CircularGrid(requidistant=False)                    */

typedef struct PyrBundleGen {
    int32_t raster;             /* PyrRasterKind                                       */
    int32_t bundle;             /* PyrBundleKind                                       */
    uint32_t flags;             /* PYR_GEN_*                                           */
    int32_t reserved0;
    int64_t param;              /* see PyrRasterKind                                   */
    int64_t first;              /* raster index of ray 0 of this call (shards, chunks) */
    int64_t total;              /* points of the whole raster: first + n <= total      */
    double lin_start, lin_step, lin_stop;   /* RECT / HEX: numpy.linspace as computed on
                                   the host, x1d[i] = i * lin_step + lin_start, last
                                   element = lin_stop (bit-identical coordinates)      */
    double aux[2];              /* HEX: shift of the second lattice                    */
    double radius;              /* collimated: pupil radius; divergent: angular radius */
    double start[3];            /* startx, starty, startz                              */
    double dir[3];              /* collimated: unit direction; divergent: angles       */
    double e[3];                /* E field unless PYR_GEN_E_PERP                       */
    double n_index;             /* background index: k = n_index * d                   */
    const int64_t *rows;        /* RECT / HEX: DEVICE pointer to 2 * nrows + 1 entries:
                                   rows[r] = number of kept points ahead of lattice row
                                   r (r = 0..nrows, nrows = param, or 2 * param for
                                   HEX: second lattice after the first), then
                                   rows[nrows + 1 + r] = first kept x index of row r   */
} PyrBundleGen;

typedef struct PyrRaysIn {
    const double *x;            /* (3, n_x)  start points, global frame            */
    const double *k;            /* (3, n)    wave vectors (complex if flag)        */
    const double *e;            /* (3, n)    E field (complex if flag); NULL ->
                                   (0, 1, 0) like ray.py:71-73                     */
    const uint8_t *alive;       /* (n_x) PYR_RAY_ALIVE bit of the producing step,
                                   NULL -> all alive                               */
    int64_t ld;                 /* leading dimension of x, k, e                    */
    int64_t n_x;                /* width of x/alive; ray i reads column i % n_x
                                   (n_x = n/2 right after a split step); 0 -> n    */
    /* wavelength batch: the call carries n_waves (2..PYR_MAX_WAVES) bundles of different
     * wavelength back to back; ray i belongs to segment w = number of entries
     * wave_end[0..n_waves-2] that are <= i.  0 or 1 = one wavelength (before.n / after.n).
     * Real-valued, non-splitting sequences of homogeneous isotropic media only.       */
    int32_t n_waves;
    int32_t reserved0;
    int64_t wave_end[PYR_MAX_WAVES];
    /* non-NULL (HOST pointer, copied into the launch): the rays are generated in the
     * prologue of the trace kernel and x / k / e / alive are ignored (may be NULL).
     * Real-valued, non-splitting sequences without E recording; anything else returns
     * PYR_E_UNSUPPORTED (generate with pyr_generate_bundle and trace the arrays).       */
    const PyrBundleGen *gen;
} PyrRaysIn;

/* flags of pyr_trace */
#define PYR_F_COMPLEX 1u        /* k and E are complex128 on input and output      */
#define PYR_F_RECORD_E 2u       /* maintain E and write out_e                      */

int pyr_version(void);
const char *pyr_strerror(int code);

/* sizeof(PyrStep) / sizeof(PyrRaysIn) as compiled, so a foreign-language binding
 * can verify its struct layout before the first call. */
int64_t pyr_sizeof_step(void);
int64_t pyr_sizeof_rays_in(void);
int64_t pyr_sizeof_bundle_gen(void);

/* Number of CUDA devices visible (0 when the driver is missing). */
int pyr_device_count(void);

/*
 * Trace `n_rays` rays through steps[0..n_steps): replaces the loop body of
 * OpticalElement.seqtrace (optical_element.py:336-375) for every ray of the
 * bundle -- Material.propagate (material_isotropic.py:238-247,
 * material_grin.py:106-220), Surface.intersect (surface.py:116-135),
 * Material.refract / reflect (material_isotropic.py:163-236,
 * material_anisotropic.py:70-155).  `steps` is HOST memory (copied into the
 * launch), the rays and all outputs are DEVICE memory on the current device.
 * One persistent kernel launch per call.  Limits of one call: 40 steps and 10 steps
 * that need an auxiliary record (explicit shape, own-frame aperture, GRIN / crystal
 * medium, partial step mode; a PYR_SHAPE_COMBINATION counts one per term) --
 * PYR_E_TOOLARGE beyond; a caller continues a longer sequence from the last record
 * (x = out_x, k = out_k, alive = out_flags of the previous call's last step), which is
 * what pyr_trace_host_io and the Python engine do.
 * Output rows: when ld_out % 16 == 0 and all output pointers are 16-byte aligned the
 * records leave as TMA bulk stores that may write the row pad up to the next multiple
 * of 2 doubles of out_x / out_k / out_e and up to the next multiple of 16 BYTES of
 * out_flags -- allocate out_flags with ld_out bytes, not n.
 */
int pyr_trace(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays,
              int64_t n_rays, uint32_t flags, void *stream);

/*
 * Spot statistics of RayBundleAnalysis (analysis/ray_analysis.py:44-86) over
 * the rays whose `flags & mask` is non-zero (flags NULL = all), about a
 * reference point `shift` (HOST pointer to 3 doubles, NULL = origin; pass e.g.
 * the vertex of the image surface so the sums do not cancel catastrophically):
 *   out[0..2] = sum (x - shift), out[3] = count, out[4..6] = sum (x - shift)^2
 * ACCUMULATED into DEVICE memory `out8` (caller zeroes it), so partial sums of
 * several ranks (same shift) can be all-reduced before centroid / rms are formed.
 */
int pyr_spot_sums(const double *x, int64_t ld, const uint8_t *flags,
                  uint32_t mask, int64_t n, const double *shift, double *out8,
                  void *stream);

/*
 * One merit-function evaluation of an optimiser loop (the caller right above the path:
 * optimize/optimize.py:73-91 calls seqtrace + RayBundleAnalysis.get_rms_spot_size once
 * per function evaluation): pyr_trace plus the sums of pyr_spot_sums over the record of the
 * LAST step (out_x / out_flags of steps[n_steps - 1], which must be set; no ray doubling in
 * the call) about `shift`, left in `spot8_dev` (DEVICE memory of 8 doubles, overwritten).
 * Where the call is one launch of the conic-only kernels the trace kernel accumulates the
 * sums itself (registers over the CTA's tiles, one reduction at the end: no second pass over
 * the record); otherwise a pyr_spot_sums launch follows.
 * spot8_host != NULL: the 8 sums are also copied to HOST memory (pinned for speed) and the
 * stream is synchronised -- one call, one 64-byte read-back.  spot8_host == NULL: the call
 * stays asynchronous like pyr_trace.
 */
int pyr_trace_spot(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays,
                   int64_t n_rays, uint32_t flags, const double *shift, double *spot8_dev,
                   double *spot8_host, void *stream);

/*
 * Reference-exact GRIN propagation of a SMALL bundle (opt-in; the fused kernels integrate
 * every ray independently): IsotropicGrinMaterial.symplecticintegrator,
 * material/material_grin.py:106-213, as written -- all rays step in lock-step until every
 * ray is final (:139), rays that are already final keep moving and are invalidated if
 * they leave the boundary afterwards (:189-190), and the energy test is summed over the
 * bundle and invalidates ALL rays (:164-176).  `step`: before = the GRIN medium, shape /
 * frames = the next surface (crossing test).  x, k (real), e (may be NULL: d = k/|k|),
 * alive (may be NULL): DEVICE (3, n) / (n) arrays of leading dimension ld; out_x / out_k:
 * the frozen state in front of the surface (k = v / n), out_alive: PYR_RAY_ALIVE where
 * valid; `scratch`: DEVICE, pyr_grin_lockstep_scratch(ld) bytes; `iterations`: DEVICE
 * int32, number of integrator steps taken = history rows.  Optional history (all NULL =
 * off): row r of ray i at hist_x[(r * 3 + c) * ld + i], hist_k likewise,
 * hist_valid[r * ld + i], r < hist_rows -- the rows the reference appends (:198-205).
 * One CTA; cost grows with n / 1024: for bundle sizes the reference itself can handle.
 */
int64_t pyr_grin_lockstep_scratch(int64_t ld);
int pyr_grin_lockstep(const PyrStep *step, const double *x, const double *k, const double *e,
                      const uint8_t *alive, int64_t ld, int64_t n, double *out_x,
                      double *out_k, uint8_t *out_alive, void *scratch, int32_t *iterations,
                      double *hist_x, double *hist_k, uint8_t *hist_valid, int64_t hist_rows,
                      void *stream);

/*
 * Spot-diagram points of OpticalSystemAnalysis.get_spot (analysis/
 * optical_system_analysis.py:283-303): x, y of the rays whose `flags & mask` is non-zero
 * (flags NULL = all), in the frame `frame` (HOST pointer, NULL = global coordinates; the
 * reference uses the last surface's frame), compacted into DEVICE array xy (2, ld_out):
 * xy[c * ld_out + j], j < *count.  *count (DEVICE int64) receives the number of selected
 * rays -- nothing is read back to the host, so the call can be followed directly by a
 * collective on fixed-width buffers.  Points beyond ld_out are counted but not stored.
 * The order of the points is unspecified (a spot diagram is a point set).
 */
int pyr_spot_points(const double *x, int64_t ld, const uint8_t *flags, uint32_t mask,
                    int64_t n, const PyrFrame *frame, double *xy, int64_t ld_out,
                    int64_t *count, void *stream);

/*
 * Write rays [0, n) of the generated bundle (raster points gen->first ...) to DEVICE
 * arrays x, k, e of leading dimension ld (any of them may be NULL): what
 * collimated_bundle / divergent_bundle return, resident on the device.
 */
int pyr_generate_bundle(const PyrBundleGen *gen, int64_t n_rays, double *x, double *k,
                        double *e, int64_t ld, void *stream);

/*
 * End-to-end host entry: start points / wave vectors / E in (pinned) HOST
 * memory, final-surface record back in HOST memory.  Chunks the bundle,
 * overlaps H2D, trace and D2H on internal streams.  `workspace` is caller-owned
 * DEVICE memory of at least pyr_trace_host_workspace() bytes.  Only real
 * (non-complex), non-splitting sequences.  Outputs (host, ld = n_rays):
 *   x_last (3, n), k_last (3, n), flags_last (n), spot8[8] (sums as above, about
 *   the origin of the last step's shape frame).
 */
int64_t pyr_trace_host_workspace(int32_t n_steps, int64_t chunk_rays);
int pyr_trace_host(const PyrStep *steps, int32_t n_steps, const double *x0,
                   const double *k0, const double *e0, int64_t n_rays,
                   double *x_last, double *k_last, uint8_t *flags_last,
                   double *spot8, void *workspace, int64_t workspace_bytes,
                   int64_t chunk_rays);

/*
 * General end-to-end host entry (pyr_trace_host is the special case "host arrays in,
 * last record out").  Inputs: host arrays x0, k0, e0 of leading dimension n_rays, or a
 * generator (no host->device traffic at all beyond the descriptor).  Outputs, all HOST
 * memory and optional: the last record (x_last, k_last, flags_last: ld = n_rays), every
 * record of the sequence like the S + 2 bundles OpticalSystem.seqtrace returns
 * (raytracer/optical_system.py:73-94, ray.py:207-260): x_all / k_all as
 * (n_steps, 3, n_rays), flags_all as (n_steps, n_rays), and the spot sums.  Sequences
 * longer than one launch are continued internally.
 * Sequences with birefringent media: real x0, k0, e0 in; the last record is that of the
 * doubled bundle (reference hstack order, material_anisotropic.py:89-100): x_last
 * (3, n m_x) and flags_last (n m_x), k_last and e_last (3, n m_k) complex128 interleaved,
 * m_x / m_k from pyr_trace_host_crystal_workspace(), which also sizes the workspace; no
 * `*_all` outputs there.
 */
typedef struct PyrHostIO {
    const double *x0, *k0, *e0;     /* e0 NULL = (0, 1, 0), ray.py:71-73                */
    const PyrBundleGen *gen;        /* non-NULL: x0 / k0 / e0 are ignored               */
    double *x_last, *k_last;
    uint8_t *flags_last;
    double *x_all, *k_all;
    uint8_t *flags_all;
    double *spot8;
    double *e_last;                 /* E of the last record (crystal sequences; complex128)     */
} PyrHostIO;

int64_t pyr_trace_host_io_workspace(int32_t n_steps, int64_t chunk_rays, int32_t all_records);
int64_t pyr_trace_host_crystal_workspace(const PyrStep *steps, int32_t n_steps, int64_t chunk_rays,
                                         int64_t *mult_x, int64_t *mult_k);
int pyr_trace_host_io(const PyrStep *steps, int32_t n_steps, const PyrHostIO *io,
                      int64_t n_rays, void *workspace, int64_t workspace_bytes,
                      int64_t chunk_rays);

#ifdef __cplusplus
}
#endif
#endif /* PYRATE_B200_H */
