/*
 * fsk_dev.h — structures shared by the host side (fsk_host.cu) and the demodulator
 * kernel (fsk_kernel.cu).  Not part of the public C-ABI.
 */
#ifndef PIRIP_FSK_DEV_H
#define PIRIP_FSK_DEV_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pirip_fsk.h"

#define FSK_MAX_FFT_FACTORS 8
#define FSK_MAX_MASK 12
#define FSK_MAX_G 8 /* captures demodulated in lockstep by one CTA */

/* Everything the kernel needs to know about one modem configuration.  Passed by
   value as a __grid_constant__ kernel parameter. */
struct FskDevCfg {
    int Fs, Rs, M, P, Nsym, Ts, N, Nmem, Ndft, Nbits, nint;
    int st, en, f_zero; /* frequency-estimator search window / blanking, in bins */
    int est_type, tone_spacing, f1_tx;
    int burst_mode, lock_nin;
    int fmt, soft;
    int nmask, len_mask;
    int mask_pos[FSK_MAX_MASK]; /* non-zero positions of the tone mask, ascending */
    int nfac;                   /* FFT factor list, outermost first (kiss_fft order) */
    int fac_p[FSK_MAX_FFT_FACTORS], fac_m[FSK_MAX_FFT_FACTORS], fac_fstride[FSK_MAX_FFT_FACTORS];
    int fac_log2m[FSK_MAX_FFT_FACTORS];
    float tc;
    float fs_over_ndft;
    const float *hann;     /* [Ndft] window, device */
    const float2 *twiddle; /* [Ndft] FFT twiddles, device */
    const float2 *phi_ft;  /* [nint] timing-line oscillator, device */
    const unsigned short *perm; /* [Ndft] swizzled FFT input permutation, device */
    int log2_ndft;
    /* launch geometry and dynamic shared memory layout (byte offsets) */
    int G;               /* captures per CTA */
    int threads;         /* threads per CTA */
    int chunk;           /* samples per pipeline chunk of the streaming phase */
    int fdc_tone_stride; /* float2 elements between the tones of one capture's f_dc */
    int cap_stride;      /* bytes per capture block */
    int off_fdc, off_xin, off_mag, off_Sf, off_stage, off_phead, off_scal; /* inside a capture block */
    int off_fft, off_tw, off_hann, off_phift, off_perm, off_ctl; /* CTA-wide, after the G capture blocks */
    int stage_bytes;
    int max_numffts;
    int mag_blocks;      /* FFT blocks per batch = capacity of the magnitude buffer */
    int pipeline;        /* the FFT block count is the same for every nin: estimate one frame ahead */
    int numffts_const;
    int smem_bytes;
    /* ---- geometry of the batch kernel (fsk_kernel2.cu): one persistent CTA per SM ---- */
    struct K2 {
        int ok;            /* configuration supported by the batch kernel */
        int gmax;          /* captures demodulated in lockstep by one CTA */
        int threads;       /* threads per CTA */
        int chunk;         /* samples per pipeline chunk of the streaming phase */
        int ring;          /* f_dc ring: elements per tone (power of two, 4 chunks) */
        int ering;         /* window-energy ring: floats per tone (power of two) */
        int ckpt_every;    /* oscillator phase checkpoint spacing (samples) */
        int nckpt;         /* checkpoints per tone */
        int tail;          /* f_dc samples kept from the end of the previous frame, per tone */
        int stage_bytes;   /* one staging buffer (two per capture) */
        int numffts;       /* FFT blocks per frame (independent of nin) */
        int fast_ts;       /* 1: the 16-window integrator path applies (P == Ts, Ts instantiated) */
        int fft_warps;     /* warps that take FFT tasks */
        int fft_early;     /* FFT tasks of the next frame a capture-less worker warp runs during the streaming phase */
        int task_queue;    /* 1: FFT tasks, per-frame sums and tone picking are handed out from one queue after the decisions */
        int ering_fft_off; /* float offset of the FFT scratch borrowed from a capture's energy ring */
        int cap_stride;    /* bytes per capture block */
        int off_stage, off_ring, off_tail, off_ckpt, off_ering, off_Sf, off_scal; /* inside a capture block */
        int off_fft, off_tw, off_hann, off_perm, off_phift, off_ctl;             /* CTA-wide, after the capture blocks */
        int smem_bytes;
        int num_sms;
    } k2;
};

/* per-capture scalars kept in shared memory for the life of the CTA */
struct FskCapShared {
    float2 phi_c[4];
    float2 dphi[2][4]; /* estimate slots: frame being demodulated / frame being estimated ahead */
    float f_est[2][4];
    float f2_est[2][4];
    int freqi[4];
    const unsigned char *fsrc; /* raw samples the front stage reads */
    int fon, fnumffts;
    float norm_rx_timing, ppm, snr_est, EbNodB;
    float fract, one_minus_fract;
    int low, high;
    int nin, nin_next, nold, numffts;
    int nan_flag, active;
    float SNRest, v_est, rx_sig_pow, rx_nse_pow, foff, rx_timing;
    float sums[4];
    unsigned long long pos;
    unsigned long long stg_lo; /* global address the staged window starts at */
    unsigned stg_bytes;
    int stg_valid;
    int frame;
    int pad;
};

/* CTA-wide control words of the streaming phase */
#define FSK_MAX_CHUNKS 64
struct FskCtaShared {
    unsigned long long bar_osc[FSK_MAX_CHUNKS]; /* mbarrier: oscillator warp finished chunk ck */
    unsigned long long bar_int[FSK_MAX_CHUNKS]; /* mbarrier: every worker warp integrated chunk ck */
    unsigned long long bar_in[FSK_MAX_G];       /* mbarrier: TMA copy of capture g's next frame landed */
    int any_active;
    int max_nin;
};

/* demodulator state carried between frames — the serial state of SURVEY.md 3.1.
   Layout in global memory: this header, then Sf[Ndft], then f_dc[M*Nmem] (float2). */
struct FskDevStateHdr {
    float2 phi_c[4];
    float norm_rx_timing, ppm, snr_est, EbNodB;
    int nin;
    int pad[3];
};

struct FskBatchArgs {
    const unsigned char *iq;     /* device */
    size_t capture_stride_bytes; /* between captures */
    size_t samples_per_capture;
    size_t max_frames;
    size_t n_captures;
    unsigned char *bits;              /* [ncap][max_frames*Nbits] or NULL */
    float *llr;                       /* [ncap][max_frames*Nbits] or NULL */
    float *rx_filt;                   /* [ncap][max_frames*M*Nsym] or NULL */
    struct pirip_frame_stats *stats;  /* [ncap][max_frames] or NULL */
    int32_t *nframes;                 /* [ncap] or NULL */
    float *state;                     /* per-capture persistent state or NULL (fresh start) */
    size_t state_stride_floats;
    float *Sf_out;                    /* optional [ncap][Ndft] copy of Sf after the last frame */
    int tma_ok;                       /* iq is 16-byte aligned: stage frames with TMA bulk copies */
    const unsigned long long *landed; /* optional (batch kernel): samples of every capture already in device memory; the
                                         host raises it while it streams the captures in, the kernel never reads ahead */
    /* optional (batch kernel): demodulate only the captures cap_list[0 .. *cap_count) of the batch (device memory, read
       when the kernel starts: the re-run of captures the fast kernel flagged); n_captures is then the launch bound and
       total_captures the size of the batch the buffers are laid out for */
    const int *cap_list;
    const int *cap_count;
    size_t total_captures;
};

static inline size_t fsk_state_floats(int Ndft, int M, int Nmem) {
    return sizeof(FskDevStateHdr) / sizeof(float) + (size_t)Ndft + 2u * (size_t)M * (size_t)Nmem;
}

/* per-capture scalars of the batch kernel (fsk_kernel2.cu) */
struct Fsk2CapShared {
    float2 phi_c[4];
    float2 dphi[2][4]; /* estimate slots by frame parity */
    float f_est[2][4];
    float f2_est[2][4];
    int freqi[4];
    float norm_rx_timing, ppm, snr_est, EbNodB;
    float fract, one_minus_fract;
    int low, high;
    int nin, nin_next, nold;
    int nan_flag, active, fon;
    float SNRest, v_est, rx_sig_pow, rx_nse_pow, foff, rx_timing;
    float sums[4];
    unsigned long long pos;
    unsigned long long stg_lo[2]; /* global address each staged window starts at */
    unsigned stg_bytes[2];
    int stg_valid[2];
    int stg_fills[2]; /* TMA copies issued into each buffer so far: phase parity of its mbarrier */
    int fast;         /* this frame takes the 16-window integrator path */
    int sf_turn;      /* FFT block whose magnitudes the IIR spectrum takes next */
    int est_done;     /* 1 + index of the last frame whose spectrum is complete (release / acquire) */
    int frame;
    int cap; /* capture index in the batch */
};

#define FSK2_MAX_CHUNKS 32
/* FFT scratch slot of one FFT warp, in float2 elements: Ndft, plus the bank skew of the 256-point exchange */
#define FSK2_FFT_SLOT_ELEMS(ndft) ((ndft) + 16)
#define FSK2_CKPT_EVERY 16 /* oscillator phase checkpoint spacing (samples) */
#define FSK2_MAX_G 16
#define FSK2_MAX_Q 640 /* post-decision queue entries: FFT tasks + 2 per capture */
#define FSK2_Q_FFT 0
#define FSK2_Q_SUMS 1
#define FSK2_Q_PICK 2
struct Fsk2CtaShared {
    unsigned long long bar_int[FSK2_MAX_CHUNKS]; /* every worker warp integrated its share of chunk c */
    unsigned long long bar_tim[FSK2_MAX_CHUNKS]; /* timing warp consumed the window energies of chunk c */
    unsigned long long bar_in[FSK2_MAX_G][2];    /* TMA copy into staging buffer b of capture g landed */
    int any_active;
    int max_nin;
    int G;
    int any_generic; /* some active capture takes the generic integrator path this frame */
    int task_next;   /* post-decision task queue: next entry to hand out */
    int decide_done; /* warps that have finished their decision lanes this frame */
    int q_len;       /* entries of the queue (fixed per capture group) */
    unsigned short queue[FSK2_MAX_Q]; /* task kind << 12 | FFT task / capture index, in hand-out order */
};

/* defined in fsk_kernel2.cu */
cudaError_t fsk2_launch_demod(const FskDevCfg &cfg, const FskBatchArgs &args, size_t n_captures, cudaStream_t stream);
cudaError_t fsk2_kernel_prepare(const FskDevCfg &cfg);

/* defined in fsk_fast.cu: the tolerance-mode batch kernel (exact tone estimates, fast integrators, guarded decisions) */
struct FskFastPlan;
#ifdef __cplusplus
#include <string>
FskFastPlan *fsk_fast_plan_create(const pirip_fsk_cfg &c, const FskDevCfg &d, const float *h_hann, const float2 *h_phi_ft, int device,
                                  std::string &why);
void fsk_fast_plan_destroy(FskFastPlan *p);
const char *fsk_fast_plan_desc(const FskFastPlan *p);
cudaError_t fsk_fast_flag_buffers(FskFastPlan *p, size_t n, int **flags, int **list, int **count);
cudaError_t fsk_fast_launch(FskFastPlan *p, const FskBatchArgs &b, int *d_flags, int *d_list, int *d_count, cudaStream_t stream,
                            size_t off, size_t total, size_t slot);
cudaError_t fsk_fast_replay_counts(FskFastPlan *p, int out[64]);
int fsk_fast_uses_chain(const FskFastPlan *p);
size_t fsk_fast_round_captures(const FskFastPlan *p);
#endif

/* defined in fsk_kernel.cu */
cudaError_t fsk_launch_demod(const FskDevCfg &cfg, const FskBatchArgs &args, size_t n_captures, cudaStream_t stream);
cudaError_t fsk_kernel_prepare(const FskDevCfg &cfg);

#endif
