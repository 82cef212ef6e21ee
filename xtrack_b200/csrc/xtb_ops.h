/* xtb_ops.h -- format of the lowered lattice ("program") that the tracking
 * kernel interprets.  Mirrored by xtrack_b200/lowering.py (single source of the
 * numeric values: keep both in sync; tests/test_lowering_hostsim.py checks it).
 *
 * The program is a stream of 8-byte words.  One op =
 *
 *   word 0   header   bits  0..7   opcode
 *                           8..15  flags   (XTB_F_*)
 *                          16..31  nwords  (length of the op in words, header included)
 *                          32..63  aux     (int32: order, counts, table index ...)
 *   word 1   L        length of the DRIFT PREFIX (valid iff XTB_F_DRIFT, else 0.0)
 *   word 2.. parameters (IEEE doubles unless stated)
 *
 * nwords is always even, so every op and its parameter block start on a
 * 16-byte boundary (bulk-copy granularity, 128-bit shared-memory loads).
 *
 * Drift prefix.  In a thin lattice every other element is a Drift; an op whose
 * header carries XTB_F_DRIFT (fast ops: also opcode bit XTB_OPBIT_DRIFT, so that one
 * indirect branch dispatches both) first performs the expanded drift of the Drift
 * ELEMENT that precedes it (track_drift.h:11-22), with that element's own
 * end-of-element actions (global aperture check, loss check, at_element + 1;
 * xtrack/tracker.py:681-711), and then its own work: two elements per
 * dispatch.
 *
 * Two programs are lowered per line:
 *   FUSED  whole single-op elements use the "fast" opcodes (< XTB_GENERIC_FIRST;
 *          implicit START|END, no GLOBAL, no monitor hook) and absorb a
 *          preceding Drift as prefix (a Drift with nothing to lead into is XTB_OP_FDRIFT);
 *   PLAIN  one element = one or more generic ops, never fused: used for
 *          launches whose element range does not fall on op boundaries of the
 *          fused program and for the element-by-element monitor.
 * Zero-coefficient specialisation (MULTP1, MULTPN, MULTH0N).  Real lattices are made of
 * plain normal magnets: one non-zero coefficient.  The reference's Horner loop
 * (track_magnet_kick.h:183-228) still multiplies and adds the zeros; the specialised ops
 * leave out exactly the operations whose operand is a literal zero coefficient
 * (0*y, z - 0, 0 + z, p + 0).  For finite coordinates these are IEEE identities, so the
 * results are the reference's bit for bit, except that a result that is exactly zero may
 * come out as -0.0 where the reference has +0.0 (or vice versa).  The same specialisation
 * exists in the reference as `SimpleThinQuadrupole` / `SimpleThinBend`
 * (simplethinquadrupole.h:13-33), which `optimize_for_tracking` swaps in (line.py:4951).
 *
 * An element that lowers to several ops carries XTB_F_START on the first and
 * XTB_F_END (loss check + at_element++) on the last, plus XTB_F_GLOBAL for
 * classes that are statically thick (global aperture check, tracker.py:681-689).
 */
#ifndef XTB_OPS_H
#define XTB_OPS_H

/* Version of the op-stream format.  Bumped whenever an opcode, a flag or a parameter
 * layout changes; xtrack_b200/lowering.py carries the same number (OPS_ABI_VERSION) and
 * _cabi.load() refuses a library whose xtb_ops_abi_version() differs: a stale libxtb200.so
 * cannot silently interpret a newer program. */
#define XTB_OPS_ABI_VERSION 7

#define XTB_F_START   0x01u
#define XTB_F_END     0x02u
#define XTB_F_GLOBAL  0x04u
#define XTB_F_DRIFT   0x08u

/* -- fast set: one whole element per op (implicit START|END, never GLOBAL) -- */
#define XTB_OP_NOP            0   /* Marker, inactive elements, plain Drift (prefix only)    */
#define XTB_OP_MULT0          1   /* [cn_0, cs_0]                                            */
#define XTB_OP_MULT1          2   /* [cn_1, cs_1, cn_0, cs_0]                                */
#define XTB_OP_MULTN          3   /* aux=order>=2; [cn_o, cs_o, ..., cn_0, cs_0]             */
#define XTB_OP_MULTPN         4   /* aux=order>=2; [cn_o, 0]  ONLY the top normal coefficient  */
                                  /* is non-zero (plain sextupole, octupole ...)             */
#define XTB_OP_MULTH0         5   /* [hl, B0, cn_0, cs_0]  order 0 with curvature, no k1     */
#define XTB_OP_EDGE           6   /* [r21, r43]            track_dipole_edge_linear.h:30-39  */
#define XTB_OP_RECT           7   /* [min_x, max_x, min_y, max_y]        limitrect.h:10-38   */
#define XTB_OP_ELLIPSE        8   /* [a_squ, b_squ, a_b_squ, 0]          limitellipse.h:13   */
/* RECT, ELLIPSE: aux = (tx16 << 16) | ty16, the high 16 bits of the high words of two lengths
 * with "|x| < tx and |y| < ty => inside": an integer pre-filter in front of the exact test */
#define XTB_OP_FDRIFT         9   /* [L, 0]  a Drift element as main op (+ global check)    */
#define XTB_OP_MULTP1        10   /* [cn_1, 0]  plain normal quadrupole kick (cs_1 = c_0 = 0)  */
#define XTB_OP_MULTH0N       11   /* [hl, B0, cn_0, 0]     MULTH0 with cs_0 == 0               */
#define XTB_OP_MULTH1N       12   /* [hl, B0, B1, cn_1, cn_0, 0]  order 1 with curvature and   */
                                  /* the k1*h term, normal components only (cs_1 = cs_0 = 0)  */
#define XTB_NUM_FAST         13
#define XTB_OPBIT_DRIFT      16   /* fast opcode | 16: the same op with a drift prefix      */
#define XTB_OP_END           31   /* tile sentinel (appended by xtb_lattice_create)         */

/* -- generic set (flags honoured; thin, always compiled) -------------------- */
#define XTB_GENERIC_FIRST    32
#define XTB_OP_GNOP          32
#define XTB_OP_DRIFT         33   /* [L]                 track_drift.h:11-22                */
#define XTB_OP_DRIFT_EXACT   34   /* [L]                 track_drift.h:26-40                */
#define XTB_OP_MULT          35   /* aux=order; [cn_o, cs_o, ..., cn_0, cs_0]               */
#define XTB_OP_MULT_H        36   /* aux=order|hasB1<<8; [hl, B0, B1, 0, cn_o, cs_o, ...]   */
#define XTB_OP_CAVITY        37   /* aux=absolute_time; [V, f, harmonic, lag, phase, 0]     */
#define XTB_OP_RFMULT        38   /* aux=order; [V, f, lag, phase, 0, {knl,ksl,pn,ps,phn,phs}*(order+1)] */
#define XTB_OP_EDGE_LIN      39   /* [r21, r43]                                             */
#define XTB_OP_SROT          40   /* [sin, cos]          track_srotation.h:12               */
#define XTB_OP_XYSHIFT       41   /* [dx, dy]            track_xyshift.h                    */
#define XTB_OP_SSHIFT        42   /* [ds]                S_SHIFT, track_misalignments.h:26  */
#define XTB_OP_YROT          43   /* [sin, cos, tan]     track_yrotation.h:12               */
#define XTB_OP_XROT          44   /* [sin, cos, tan]     track_xrotation.h:12               */
#define XTB_OP_LIMIT_RECT    45   /* [min_x, max_x, min_y, max_y]                           */
#define XTB_OP_LIMIT_ELLIPSE 46   /* [a_squ, b_squ, a_b_squ]                                */
#define XTB_OP_LIMIT_POLYGON 47   /* aux=N; [x_0..x_N-1, y_0..y_N-1, (pad)]                 */
#define XTB_OP_MONITOR       48   /* aux=index into the in-line ParticlesMonitor table      */
#define XTB_OP_LAST_TURNS    49   /* aux=index into the in-line LastTurnsMonitor table      */
#define XTB_OP_KILL          50   /* aux=state code: LocalParticle_kill_particle            */
#define XTB_OP_SET_STATE     51   /* aux=state code (e.g. -42 invalid thin slice transform) */
#define XTB_OP_ADD_S_ZETA    52   /* [ds]  s += ds; zeta += ds  (track_magnet.h:598-605)    */
#define XTB_OP_ADD_X         53   /* [dx]  x += dx              (rbend straight body)       */
#define XTB_OP_BEAM_MON      54   /* aux=#sums (3 position, 5 size); see beam_monitor_record */
#define XTB_OP_BEAM_PROFILE  55   /* BeamProfileMonitor; see beam_profile_record              */
#define XTB_OP_BEAM_STATS    57   /* [(ptr) descriptor, 0]  BeamStatsMonitor; see beam_stats_record */
#define XTB_OP_CRAB          56   /* aux=absolute_time; [V_t, f, lag, phase]  track_rf.h:116-156 */

/* -- heavy set (thick magnets; compiled in the HEAVY kernel variants) ------ */
#define XTB_HEAVY_FIRST      64
#define XTB_OP_MAGNET_BODY   64   /* see xtb_thick.cuh                                      */
#define XTB_OP_MAGNET_EDGE   65   /* full / dipole-only edge (track_magnet_edge.h model 1,2)*/
#define XTB_OP_DIPEDGE_NL    66   /* DipoleEdge element, model 1                            */

#define XTB_HDR(op, flags, nwords, aux) \
    ((uint64_t)(op) | ((uint64_t)(flags) << 8) | ((uint64_t)(nwords) << 16) | \
     ((uint64_t)(uint32_t)(aux) << 32))

/* element offsets handed to xtb_lattice_create: word offset of the element's
 * first op, or XTB_NOT_ADDRESSABLE when the element is absorbed in an op that
 * starts at an earlier element (main part of a drift-prefixed op) */
#define XTB_NOT_ADDRESSABLE 0xffffffffu

/* tiling of the program through shared memory */
#ifndef XTB_TILE_WORDS
#define XTB_TILE_WORDS   1024      /* 8 KiB per tile (>= TILE_WORDS of lowering.py)   */
#endif
#define XTB_NUM_BUF      2

#endif /* XTB_OPS_H */
