/* xtb_ops.h -- format of the lowered lattice ("program") that the tracking
 * kernel interprets.  Mirrored by xtrack_b200/lowering.py (single source of the
 * numeric values: keep both in sync; tests/test_lowering.py checks it).
 *
 * The program is a stream of 8-byte words.  One op = one header word followed
 * by `nwords-1` parameter words (IEEE doubles unless stated); nwords is always
 * even, so every op starts on a 16-byte boundary (bulk-copy granularity).
 *
 *   header bits  0..7   opcode
 *                8..15  flags   (XTB_F_*)
 *               16..31  nwords  (length of the op in words, header included)
 *               32..63  aux     (int32: order, counts, table index ...)
 *
 * An xtrack element lowers to one or more ops; the first carries
 * XTB_F_START, the last XTB_F_END (loss check + at_element++,
 * xtrack/tracker.py:702-711) and, for classes that are statically thick,
 * XTB_F_GLOBAL (global aperture check, tracker.py:681-689).
 */
#ifndef XTB_OPS_H
#define XTB_OPS_H

#define XTB_F_START   0x01u
#define XTB_F_END     0x02u
#define XTB_F_GLOBAL  0x04u

/* -- thin set (always compiled) ------------------------------------------ */
#define XTB_OP_NOP            0   /* Marker, inactive elements                              */
#define XTB_OP_DRIFT          1   /* [L]                 track_drift.h:11-22                */
#define XTB_OP_DRIFT_EXACT    2   /* [L]                 track_drift.h:26-40                */
#define XTB_OP_MULT           3   /* aux=order; [cn_o, cs_o, ..., cn_0, cs_0]               */
#define XTB_OP_MULT_H         4   /* aux=order|hasB1<<8; [hl, B0, B1, 0, cn_o, cs_o, ...]   */
#define XTB_OP_CAVITY         5   /* aux=absolute_time; [V, f, harmonic, lag, phase, 0, 0]  */
#define XTB_OP_RFMULT         6   /* aux=order; [V, f, lag, phase, 0, {knl,ksl,pn,ps,phn,phs}*(order+1)] */
#define XTB_OP_EDGE_LIN       7   /* [r21, r43, 0]       track_dipole_edge_linear.h:30-39   */
#define XTB_OP_SROT           8   /* [sin, cos, 0]       track_srotation.h:12               */
#define XTB_OP_XYSHIFT        9   /* [dx, dy, 0]         track_xyshift.h                    */
#define XTB_OP_SSHIFT        10   /* [ds]                S_SHIFT, track_misalignments.h:26  */
#define XTB_OP_YROT          11   /* [sin, cos, tan]     track_yrotation.h:12               */
#define XTB_OP_XROT          12   /* [sin, cos, tan]     track_xrotation.h:12               */
#define XTB_OP_LIMIT_RECT    13   /* [min_x, max_x, min_y, max_y, 0]                        */
#define XTB_OP_LIMIT_ELLIPSE 14   /* [a_squ, b_squ, a_b_squ]                                */
#define XTB_OP_LIMIT_POLYGON 15   /* aux=N; [x_0..x_N-1, y_0..y_N-1, (pad)]                 */
#define XTB_OP_MONITOR       16   /* aux=index into the in-line ParticlesMonitor table      */
#define XTB_OP_LAST_TURNS    17   /* aux=index into the in-line LastTurnsMonitor table      */
#define XTB_OP_KILL          18   /* aux=state code: LocalParticle_kill_particle            */
#define XTB_OP_SET_STATE     19   /* aux=state code (e.g. -42 invalid thin slice transform) */
#define XTB_OP_ADD_S_ZETA    20   /* [ds]  s += ds; zeta += ds  (track_magnet.h:598-605)    */
#define XTB_OP_ADD_X         21   /* [dx]  x += dx              (rbend straight body)       */

/* -- heavy set (thick magnets; compiled in the HEAVY kernel variants) ------ */
#define XTB_OP_MAGNET_BODY   32   /* see xtb_thick.cuh                                      */
#define XTB_OP_MAGNET_EDGE   33   /* full / dipole-only edge (track_magnet_edge.h model 1,2)*/
#define XTB_OP_DIPEDGE_NL    34   /* DipoleEdge element, model 1                            */
#define XTB_OP_RF_BODY       35   /* thick cavity / rf-multipole body                       */

#define XTB_HEAVY_FIRST      32

#define XTB_HDR(op, flags, nwords, aux) \
    ((uint64_t)(op) | ((uint64_t)(flags) << 8) | ((uint64_t)(nwords) << 16) | \
     ((uint64_t)(uint32_t)(aux) << 32))

/* tiling of the program through shared memory */
#define XTB_TILE_WORDS   1024      /* 8 KiB per tile                                  */
#define XTB_NUM_BUF      2

#endif /* XTB_OPS_H */
