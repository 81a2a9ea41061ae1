"""Single source of truth for the device ModelBlob and the env-minor Data buffer layout.

``python -m mujoco_torch_b200.layout`` regenerates ``csrc/mjt_layout.h`` (committed), so the CUDA
kernels and the Python packer agree on every slot index.

ModelBlob (one contiguous int32 buffer, copied to shared memory once per CTA with a TMA bulk copy):
    [ H scalars | I offsets | F offsets | D offsets | D sizes | P offsets | int arrays ... | pad16 | float arrays ... ]
Float arrays are stored in the compute dtype (f32 or f64); F offsets are in elements from the float base.
The Data buffer is a (W, B) array, env-minor: word k of env e lives at buf[k * B + e].
"""

from __future__ import annotations

import os

# ---- header scalars (int32) ---------------------------------------------------------------
H_SCALARS = """
magic blob_words float_base_words is_f64
nq nv nu na nbody njnt ngeom nsite nmocap ntendon nwrap neq nsensor nsensordata
ncon nefc ne nf nl npair_cand nsrc
disableflags solver cone integrator iterations ls_iterations
data_words has_gravcomp has_fluid nv_pad nlim_sh nlim_ball nfric_dof nfric_ten nlim_ten
ne_connect ne_weld ne_joint maskwords nsens_pos nsens_vel nsens_acc
nconvex nrfgeom jc_words solve_blob_words live_rcap live_jcap
layout_state pub_words n_mnz topk ncand nsrc1
""".split()

# ---- int arrays ---------------------------------------------------------------------------
I_ARRAYS = """
body_parentid body_rootid body_jntadr body_jntnum body_dofadr body_dofnum body_mocapid
jnt_type jnt_qposadr jnt_dofadr jnt_bodyid jnt_limited
dof_bodyid dof_parentid dof_jntid
geom_bodyid geom_type site_bodyid
body_dofmask
act_trntype act_trnid act_dyntype act_gaintype act_biastype act_actadr act_actnum
act_ctrllimited act_forcelimited act_actlimited
jnt_actfrclimited jnt_actgravcomp
ten_adr ten_num wrap_jnt ten_limited
pair_fn pair_g1 pair_g2 pair_slot
con_body1 con_body2 con_dim con_efcadr con_geom1 con_geom2
src_type src_id src_row src_cadr src_nsup src_nrows sup_col row_src row_jadr jc_src tri_src triu_src tri_ab dof_sadr dof_sent
eq_type eq_obj1 eq_obj2
sens_type sens_objid sens_adr sens_dim sens_datatype sens_stage sens_objtype sens_reftype sens_refid
rf_geom
body_geomadr body_geomnum body_geom body_siteadr body_sitenum body_site
cvx_geom cvx_nvert cvx_nface cvx_nfv cvx_nedge cvx_vertadr cvx_faceadr cvx_edgeadr cvx_normadr geom_cvx cvx_face cvx_edge
m_nz geom_need row_order dof_ent topk_perm
hf_nrow hf_ncol hf_adr pair_hf
""".split()

# ---- float arrays -------------------------------------------------------------------------
F_ARRAYS = """
opt
qpos0 qpos_spring
body_pos body_quat body_ipos body_iquat body_mass body_subtreemass body_inertia body_gravcomp body_invweight0
jnt_pos jnt_axis jnt_stiffness jnt_range jnt_margin jnt_solref jnt_solimp jnt_actfrcrange
dof_armature dof_damping dof_invweight0 dof_frictionloss dof_solref dof_solimp
geom_pos geom_quat geom_size site_pos site_quat
act_gear act_ctrlrange act_forcerange act_actrange act_gainprm act_biasprm act_dynprm act_lengthrange act_acc0
wrap_prm ten_stiffness ten_damping ten_armature ten_lengthspring ten_range ten_margin ten_invweight0
ten_solref_lim ten_solimp_lim ten_frictionloss ten_solref_fri ten_solimp_fri
con_includemargin con_friction con_solref con_solreffriction con_solimp
eq_data eq_solref eq_solimp
sens_cutoff
cvx_vert cvx_norm
hf_size hf_data pair_rbound
""".split()

# slots inside the ``opt`` float array
OPT_SLOTS = "timestep impratio tolerance ls_tolerance meaninertia gravity0 gravity1 gravity2 wind0 wind1 wind2 density viscosity magnetic0 magnetic1 magnetic2".split()

# ---- Data fields: name -> per-env shape as a python expression of model sizes ---------------
# (order matters only for readability; offsets are computed at device_put time)
D_FIELDS = [
    # state / integrator outputs
    ("qpos", "(nq,)"), ("qvel", "(nv,)"), ("act", "(na,)"), ("time", "()"), ("qacc", "(nv,)"),
    ("qacc_warmstart", "(nv,)"), ("act_dot", "(na,)"),
    # kinematics
    ("xpos", "(nbody,3)"), ("xquat", "(nbody,4)"), ("xmat", "(nbody,3,3)"), ("xipos", "(nbody,3)"),
    ("ximat", "(nbody,3,3)"), ("xanchor", "(njnt,3)"), ("xaxis", "(njnt,3)"),
    ("geom_xpos", "(ngeom,3)"), ("geom_xmat", "(ngeom,3,3)"), ("site_xpos", "(nsite,3)"), ("site_xmat", "(nsite,3,3)"),
    # com / inertia
    ("subtree_com", "(nbody,3)"), ("cinert", "(nbody,10)"), ("crb", "(nbody,10)"), ("cdof", "(nv,6)"),
    ("cdof_dot", "(nv,6)"), ("cvel", "(nbody,6)"), ("qM", "(nv,nv)"), ("qLD", "(nv,nv)"),
    # tendons / actuators
    ("ten_length", "(ntendon,)"), ("ten_J", "(ntendon,nv)"), ("ten_velocity", "(ntendon,)"),
    ("actuator_length", "(nu,)"), ("actuator_moment", "(nu,nv)"), ("actuator_velocity", "(nu,)"), ("actuator_force", "(nu,)"),
    # forces
    ("qfrc_bias", "(nv,)"), ("qfrc_passive", "(nv,)"), ("qfrc_gravcomp", "(nv,)"), ("qfrc_actuator", "(nv,)"),
    ("qfrc_smooth", "(nv,)"), ("qacc_smooth", "(nv,)"), ("qfrc_constraint", "(nv,)"),
    # contacts / constraints
    ("contact_dist", "(ncon,)"), ("contact_pos", "(ncon,3)"), ("contact_frame", "(ncon,3,3)"),
    ("efc_J", "(nefc,nv)"), ("efc_D", "(nefc,)"), ("efc_aref", "(nefc,)"), ("efc_frictionloss", "(nefc,)"), ("efc_force", "(nefc,)"),
    ("sensordata", "(nsensordata,)"),
    # scratch used between kernels (not exposed)
    ("scr_cacc", "(nbody,6)"), ("scr_cfrc", "(nbody,6)"),
    # per row source: 1 when the constraint kernel instantiated non-zero rows for it this step; per env: 1 when the live
    # tile of the solver overflowed and the env has to be solved by the full-tile kernel
    ("scr_live", "(nsrc,)"), ("scr_ovf", "()"),
    # max_contact_points (top-k contact selection, reference collision_driver.py:822-840): every candidate contact in group order,
    # and per final contact slot the candidate it came from (all of size 0 for models without the numeric)
    ("scr_cand_dist", "(ncand,)"), ("scr_cand_pos", "(ncand,3)"), ("scr_cand_frame", "(ncand,3,3)"), ("scr_consrc", "(ntopk,)"),
]

# kernel dispatch "source" types for the constraint kernel
SRC_TYPES = "EQ_CONNECT EQ_WELD EQ_JOINT FRIC_DOF FRIC_TEN LIM_BALL LIM_SH LIM_TEN CONTACT".split()
# narrowphase function ids
PAIR_FNS = ("PLANE_SPHERE PLANE_CAPSULE SPHERE_SPHERE SPHERE_CAPSULE CAPSULE_CAPSULE PLANE_CONVEX SPHERE_CONVEX CAPSULE_CONVEX CONVEX_CONVEX "
            "HFIELD_SPHERE HFIELD_CAPSULE HFIELD_CONVEX").split()
PAIR_NCON = dict(PLANE_SPHERE=1, PLANE_CAPSULE=2, SPHERE_SPHERE=1, SPHERE_CAPSULE=1, CAPSULE_CAPSULE=1, PLANE_CONVEX=4,
                 SPHERE_CONVEX=1, CAPSULE_CONVEX=2, CONVEX_CONVEX=4, HFIELD_SPHERE=4, HFIELD_CAPSULE=4, HFIELD_CONVEX=4)

HIDDEN_FIELDS = {"scr_cacc", "scr_cfrc", "scr_live", "scr_ovf", "scr_cand_dist", "scr_cand_pos", "scr_cand_frame", "scr_consrc"}

# ---- the "state" output layout (``step(..., outputs="state")``) --------------------------------------------------
# Same kernels, second blob: its D-offset table describes a split buffer.  Words [0, pub_words) are the PUBLIC part -- the
# fields below, a small (pub_words, B) buffer returned to the caller; words >= pub_words live in a WORKSPACE the caller
# lends to the step (scratch, reused from step to step).  Fields nothing downstream reads (xmat, cinert, cdof_dot, dense qM /
# qLD / efc_J / actuator_moment, qacc_smooth, efc_force ...) get offset -1 and are never written; qM is stored as a packed
# lower triangle and efc_J compacted to each row's column support (the form k_solve keeps in shared memory anyway).
# (the state a rollout feeds back in -- qpos, qvel, act, time, qacc_warmstart -- comes first and is contiguous, so a host pipeline
# moves it with one copy per direction)
STATE_PUBLIC = ("qpos", "qvel", "act", "time", "qacc_warmstart", "qacc", "act_dot", "sensordata", "qfrc_constraint")
# P table of the blob header: word offsets of the public fields inside the buffer handed back to the caller.  Full layout:
# the same buffer and offsets as the D table.  State layout: offsets inside the small public buffer (-1 = absent); the D
# table then addresses the workspace, which keeps its own pre-integration working copies of qpos / qvel / act / time.
P_FIELDS = STATE_PUBLIC
MAGIC = 0x4D4A5431  # "MJT1"


def h_index(name):
    return H_SCALARS.index(name)


def header_text() -> str:
    out = ["// GENERATED by mujoco_torch_b200/layout.py -- do not edit.", "#pragma once", ""]

    def enum(prefix, names, count_name):
        out.append("enum {")
        for i, n in enumerate(names):
            out.append(f"  {prefix}{n} = {i},")
        out.append(f"  {count_name} = {len(names)}")
        out.append("};")
        out.append("")

    enum("H_", H_SCALARS, "NUM_H")
    enum("I_", I_ARRAYS, "NUM_I")
    enum("F_", F_ARRAYS, "NUM_F")
    enum("D_", [n for n, _ in D_FIELDS], "NUM_D")
    enum("P_", list(P_FIELDS), "NUM_P")
    enum("OPT_", OPT_SLOTS, "NUM_OPT")
    enum("SRC_", SRC_TYPES, "NUM_SRC")
    enum("FN_", PAIR_FNS, "NUM_FN")
    out.append(f"#define MJT_MAGIC 0x{MAGIC:08X}")
    out.append("#define MJT_OFF_I (NUM_H)")
    out.append("#define MJT_OFF_F (NUM_H + NUM_I)")
    out.append("#define MJT_OFF_D (NUM_H + NUM_I + NUM_F)")
    out.append("#define MJT_OFF_DSZ (NUM_H + NUM_I + NUM_F + NUM_D)")
    out.append("#define MJT_OFF_P (NUM_H + NUM_I + NUM_F + 2 * NUM_D)")
    out.append("#define MJT_HDR_WORDS (NUM_H + NUM_I + NUM_F + 2 * NUM_D + NUM_P)")
    out.append("")
    return "\n".join(out)


def write_header():
    path = os.path.join(os.path.dirname(__file__), "csrc", "mjt_layout.h")
    txt = header_text()
    old = open(path).read() if os.path.exists(path) else None
    if old != txt:
        with open(path, "w") as fh:
            fh.write(txt)
    return path


if __name__ == "__main__":
    print(write_header())
