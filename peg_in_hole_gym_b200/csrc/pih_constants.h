/* Physics / scene constants (DATA only) shared by the CPU oracle and the CUDA kernels.
 * Every value cites the reference line it restates (paths relative to the reference repo) or is
 * marked NEW (reference silent; defined in DESIGN.md) / [3P] (PyBullet default, unverified here). */
#ifndef PIH_CONSTANTS_H
#define PIH_CONSTANTS_H

#define PIH_NB_MAX 9            /* panda: 7 arm + 2 fingers; ur5: 6 */
#define PIH_MAXC 8              /* contact points kept per env per step (deepest-first) */
#define PIH_STATE_F4 12         /* float4 slots of env state (layout: DESIGN.md section 4) */
#define PIH_STATE_WORDS (4 * PIH_STATE_F4)

/* state word indices inside the [PIH_STATE_F4][N][4] float4-SoA buffer */
#define PIH_S_Q 0               /* 0..8   joint positions */
#define PIH_S_STEP 9            /* int bits: steps taken in this episode */
#define PIH_S_EPISODE 10        /* int bits: episode index (RNG counter) */
#define PIH_S_FLAGS 11          /* int bits: bit0 done, bit1 success, bits 8..15 contact-pair mask */
#define PIH_S_QD 12             /* 12..20 joint velocities */
#define PIH_S_PPOS 24           /* 24..26 peg / free-body position (env frame) */
#define PIH_S_PQUAT 27          /* 27..30 orientation quaternion x,y,z,w */
#define PIH_S_PVEL 32           /* 32..34 linear velocity */
#define PIH_S_PANG 35           /* 35..37 angular velocity (env frame) */
/* scripted grasp task only (envs/peg_in_hole.py:254-267): */
#define PIH_S_FSM_STATE 21      /* int bits: cur_state 0..9 */
#define PIH_S_FSM_T 22          /* int bits: engine steps already spent in cur_state (state_t / timeStep) */
#define PIH_S_RANDY 23          /* random_vector[1] ~ U(-0.03, 0.03) */
#define PIH_S_MTGT 40           /* 40..46 arm joint position targets that persist between FSM states */
#define PIH_S_GRASP_ANG 47      /* grasp angle captured on entering FSM state 2 (peg_in_hole.py:65-72): labels of random_grasp */
/* scripted task, float4 slots 12..15 (words 48..63; the buffer of that task has PIH_SCRIPTED_STATE_F4 slots): the scene
 * latched when the state machine entered state 2 - the moment the reference takes `self.grasp_img = self.render()`
 * (peg_in_hole.py:65-66) - so that the image and the grasp labels describe the same instant: arm joints and the pose of
 * the grasped object (peg / grasp link).  A zero quaternion = nothing captured yet. */
#define PIH_SCRIPTED_STATE_F4 16
#define PIH_S_LATCH_Q 48        /* 9 joint positions */
#define PIH_S_LATCH_PPOS 57     /* 3 */
#define PIH_S_LATCH_PQUAT 60    /* 4 */

#define PIH_FLAG_DONE 1
#define PIH_FLAG_SUCCESS 2
#define PIH_FLAG_FAULT 4        /* NEW: a non-finite value appeared in the env's state during the step; the env
                                   was re-initialised in the same launch (new episode), reported done with reward 0 */
#define PIH_FLAG_PREV_SUCCESS 8 /* NEW, auto_reset only: the episode that ended in the step that set this word ended in success
                                   (the success bit itself is cleared together with done when the env restarts) */
#define PIH_PAIR_SHIFT 8
#define PIH_PAIR_PEG_TABLE 1    /* contact-pair presence bits (<< PIH_PAIR_SHIFT) */
#define PIH_PAIR_PEG_HOLE 2
#define PIH_PAIR_PEG_FINGER1 4
#define PIH_PAIR_PEG_FINGER2 8
#define PIH_PAIR_ARM_TABLE 16
#define PIH_PAIR_FB_ARM 32       /* free body vs arm collision spheres (hand / wrist links) */

/* tasks / arms */
#define PIH_TASK_PEG_IN_HOLE 0
#define PIH_TASK_RANDOM_FLY 1
#define PIH_TASK_SCRIPTED 2     /* the scripted grasp->insert state machine the reference ships (peg_in_hole.py:39-225) */
#define PIH_ARM_PANDA 0
#define PIH_ARM_UR5 1

/* world */
#define PIH_GRAVITY 9.8                 /* envs/peg_in_hole.py:230 gravity (0,0,-9.8) */
#define PIH_DT_DEFAULT (1.0 / 240.0)    /* envs/peg_in_hole.py:254; Bullet default [3P] */
#define PIH_SOLVER_ITERS 50             /* Bullet default numSolverIterations [3P] */
#define PIH_ERP 0.2                     /* Bullet default erp / contact erp [3P] */
#define PIH_CONTACT_MARGIN 0.002        /* NEW: contacts created when distance < margin */
#define PIH_MAX_COORD_VEL 100.0          /* btMultiBody m_maxCoordinateVelocity [3P]: clamp on joint rates, and on the
                                           free body's |v| (m/s) and |w| (rad/s); keeps the explicit gyroscopic term bounded */
#define PIH_TABLE_Z (-0.05)             /* table top: base z -1.3 + 2*(0.6+0.025), envs/utils.py:24-28, peg_in_hole.py:235 */

/* IK: calculateInverseKinematics defaults [3P] */
#define PIH_IK_ITERS 20
#define PIH_IK_RESIDUAL 1e-4
#define PIH_IK_LAMBDA 0.6               /* BussIK DefaultDampingLambda as shipped in Bullet [3P] */
#define PIH_IK_MAXANGLE 0.7853981633974483  /* 45 deg per-iteration clamp [3P] */
#define PIH_DV (2.0 / 240.0)            /* envs/utils.py:60 panda_execute dv */

/* peg: one rigid end segment of pipe.urdf (globalScaling 0.01): cylinder r=1,len=6 mesh units along local y
 * (assets/urdf/pipe.urdf:29-35,53-54; envs/peg_in_hole.py:242); mass/inertia as Bullet box-of-AABB [3P] */
#define PIH_PEG_R 0.01
#define PIH_PEG_HL 0.03
#define PIH_PEG_MASS 0.0111
#define PIH_PEG_MU 100.0                /* pipe.urdf:10 lateral_friction */
#define PIH_PEG_SPAWN_Z 0.11            /* envs/peg_in_hole.py:239 */
/* hole: annular tube cylinder_base.obj scaled 0.016, yaw pi/2 => axis along env x (hole.urdf, peg_in_hole.py:248-251) */
#define PIH_HOLE_X 0.5
#define PIH_HOLE_Y (-0.2)
#define PIH_HOLE_Z 0.2
#define PIH_HOLE_RIN 0.01536
#define PIH_HOLE_ROUT 0.0192
#define PIH_HOLE_HL 0.016
#define PIH_SUCCESS_RADIUS 0.05         /* envs/peg_in_hole.py:114 */
#define PIH_MU_DEFAULT 0.5              /* Bullet default lateral friction [3P] */
#define PIH_MU_FINGER 1.0               /* panda finger lateral_friction [3P] */

/* flexible pipe (assets/urdf/pipe.urdf, globalScaling 0.01, peg_in_hole.py:240-245): 25 links pipe_link0..24, each a
 * cylinder r = 1, y in [0, 6] mesh units in its link frame (collision origin (0,3,0), pipe.urdf:29-35); link1 hangs on
 * link0 by a FIXED joint at (0,3,0) (pipe.urdf:39-43), links 2..24 by continuous joints at (0,5.5,0) of their parent with
 * axes alternating z (even link numbers) / x (odd) (pipe.urdf:76-83,109-116,...).  Bodies of the multibody: 0 = link0 +
 * link1 (floating base), b = 1..23 = link b+1.  Masses pipe.urdf:16,54,87,...; link1's inertial origin (0,1.5,0)
 * (pipe.urdf:53), every other link's at its frame origin; inertia = box of the collision AABB + 1 mm margin about that
 * origin (Bullet default URDF loading [3P]); lateral friction 100 on links 0, 1, 23, 24 (pipe.urdf:10,48,765,802). */
#define PIH_PEG_RIGID 0         /* pih_config.peg_model: one rigid end segment (round-1 stand-in) */
#define PIH_PEG_PIPE 1          /* pih_config.peg_model: the 25-link chain */
#define PIH_PIPE_NSEG 25
#define PIH_PIPE_NBODY 24
#define PIH_PIPE_NJ 23
#define PIH_PIPE_NDOF 29        /* 6 base + 23 joints */
#define PIH_PIPE_R 0.01
#define PIH_PIPE_HL 0.03
#define PIH_PIPE_PITCH 0.055
#define PIH_PIPE_LINK1_Y 0.03
#define PIH_PIPE_LINK_MASS 0.0111
#define PIH_PIPE_LINK0_MASS 0.00111
#define PIH_PIPE_LINK1_COM_Y 0.015
#define PIH_PIPE_MU_END 100.0
#define PIH_PIPE_MOTOR_IMPULSE 1.0   /* default velocity motor of every movable joint: target 0, max impulse 1 per step [3P] */
#define PIH_PIPE_BEND_MAX 1.0471975511965976   /* random.uniform(0, pi/3), peg_in_hole.py:245 */
#define PIH_PIPE_SEG_MAXC 2     /* contact points kept per segment per step (deepest first) */
#define PIH_PIPE_ARM_MAXC 4     /* arm-sphere <-> table contacts kept per step (deepest first) */
#define PIH_PIPE_SEG_TOTAL 32   /* pipe contact points solved per step: every segment's deeper point first (segment order), then the
                                   second points, until 32 (a pipe lying flat keeps 25 + 7) */
#define PIH_PIPE_MAXC (PIH_PIPE_SEG_TOTAL + PIH_PIPE_ARM_MAXC + 1)   /* + the scripted task's grasp constraint */
#define PIH_PIPE_ARMROWS 16     /* of those, contact points that involve the arm (finger / arm sphere <-> segment, arm <-> table):
                                   later ones are dropped */
/* pipe task state: float4 slots 16..31 of a PIH_PIPE_STATE_F4-slot buffer (slots 0..11 as above, 12..15 reserved) */
#define PIH_PIPE_STATE_F4 32
#define PIH_PS_BASE 64          /* first word of the pipe block */
#define PIH_PS_POS (PIH_PS_BASE + 0)    /* 3: base (pipe_link0 frame) position, env frame */
#define PIH_PS_QUAT (PIH_PS_BASE + 3)   /* 4: x,y,z,w */
#define PIH_PS_VEL (PIH_PS_BASE + 7)    /* 3: linear velocity of the base origin, env frame */
#define PIH_PS_ANG (PIH_PS_BASE + 10)   /* 3: angular velocity, env frame */
#define PIH_PS_GRASP (PIH_PS_BASE + 13) /* int bits: grasp_joint_idx, 0 or 23 (peg_in_hole.py:266) */
#define PIH_PS_QJ (PIH_PS_BASE + 16)    /* 23 joint angles */
#define PIH_PS_QDJ (PIH_PS_BASE + 40)   /* 23 joint rates */
#define PIH_PAIR_GRASP 64       /* the scripted task's grasp constraint (peg_in_hole.py:99-104) is active */
/* createConstraint(panda, 11, object, grasp_joint_idx, JOINT_GEAR, [0,0,0], [0,0,0], random_vector, childFrameOrientation) on
 * entering FSM state 4, removeConstraint on entering state 7 (peg_in_hole.py:99-104).  What Bullet's gear constraint makes
 * of these arguments cannot be established ([3P]: link 11 hangs on a FIXED joint and owns no degree of freedom for a gear
 * row).  Built here is the evident intent - the object's grasp point (child frame position = random_vector) is pinned to the
 * end-effector origin (parent frame position [0,0,0]) while the constraint exists: three bilateral velocity rows along the
 * env axes, position error corrected with ERP, |force| <= createConstraint's default maxForce 500 N per row.  NEW. */
#define PIH_GRASP_MAX_FORCE 500.0
#define PIH_FSM_GRASP_FIRST 4
#define PIH_FSM_GRASP_LAST 6

/* random-fly free bodies (assets/urdf/banana.urdf:5-15; Amicelli_800_tex.urdf:5-15): mass 1, mu 0;
 * collision = AABB box of the convex pieces */
#define PIH_FLY_MASS 1.0
#define PIH_FLY_MU 0.0
#define PIH_BANANA_HX 0.09865
#define PIH_BANANA_HY 0.0376
#define PIH_BANANA_HZ 0.0187
#define PIH_AMICELLI_HX 0.0225
#define PIH_AMICELLI_HY 0.0212
#define PIH_AMICELLI_HZ 0.037

/* scripted grasp FSM (envs/peg_in_hole.py:122-212, 254-267) */
#define PIH_FSM_NSTATES 10
/* engine steps spent in each state: stateDurations [0.25,2,2,1,1.5,1.5,0.5,0.25,0.25,0.25] s with the reference's
 * `state_t += 1/240; if state_t > duration` double accumulation (peg_in_hole.py:206-212,263): 2226 steps in total */
#define PIH_FSM_STEPS_LIST 59, 481, 481, 241, 361, 361, 121, 60, 60, 1
#define PIH_FSM_DV 0.05                 /* self.dv, peg_in_hole.py:257 */
#define PIH_FSM_KP 0.1                  /* setJointMotorControl2 default positionGain [3P] */
#define PIH_FSM_ARM_FORCE (5.0 * 240.0) /* peg_in_hole.py:140,149,164,175,185,200 */
#define PIH_FSM_FINGER_OPEN 0.02        /* peg_in_hole.py:131,190 (force 20) */
#define PIH_FSM_FINGER_CLOSED 0.006     /* peg_in_hole.py:154 (force 20000) */
#define PIH_FSM_FINGER_FORCE_LO 20.0
#define PIH_FSM_FINGER_FORCE_HI 20000.0

/* camera of PegInHole.render (envs/peg_in_hole.py:276-304): eye at the end effector looking along -z, up +y,
 * FOV 60 deg, aspect 1, near 0.001, far 1000, 300 x 300; output channels (depth, r, g, b) */
#define PIH_LABEL_LENGTH 0.1     /* peg_in_hole.py:73-74: half extents of the grasp rectangle in normalised image units */
#define PIH_LABEL_WIDTH 0.2
#define PIH_LABEL_POS_VALUE 50.0 /* peg_in_hole.py:87 */
#define PIH_LABEL_NONE (-100.0)  /* grasp angle word before FSM state 2 has been entered: the label images are still empty */
#define PIH_CAM_W 300
#define PIH_CAM_H 300
#define PIH_CAM_TAN_HALF_FOV 0.5773502691896257   /* tan(30 deg) */
#define PIH_CAM_NEAR 0.001
#define PIH_CAM_FAR 1000.0
#define PIH_SCENE_WORDS 64      /* per-env scene block handed from the scene pre-pass to the ray caster */

/* RNG stream ids (Philox4x32-10 counter word 2) */
#define PIH_STREAM_RESET 0
#define PIH_STREAM_ACTION 1     /* (+1 for the second half of a Box(6) action) */
#define PIH_STREAM_FLY 8        /* 8 + k: attempt k of the rejection loop of random_pos_in_panda_space (envs/utils.py:97-107) */
#define PIH_FLY_MAX_ATTEMPTS 8  /* NEW bound on that loop (a rejection has probability ~3 %: 8 in a row never happens in practice) */
#define PIH_STREAM_PIPE 16      /* 16..: pipe bends and grasp link at reset (peg_in_hole.py:244-245,266) */

/* pih_set_stats counters (uint64[8]), accumulated by the step / rollout kernels when a buffer is set */
#define PIH_STAT_ENV_STEPS 0
#define PIH_STAT_IK_ITERS 1     /* damped-least-squares updates executed */
#define PIH_STAT_PGS_ROW_ITERS 2 /* constraint rows solved, summed over iterations (a converged island's rows do not count) */
#define PIH_STAT_PGS_ITERS 3    /* solver iterations until the env's last island converged */
#define PIH_STAT_CONTACTS 4     /* contact points (3 rows each) */
#define PIH_STAT_LIMIT_ROWS 5
#define PIH_STAT_FAULTS 6       /* envs force-reset by the non-finite guard */
#define PIH_STAT_PGS_ITERS_MAX 7

#endif
