// fse2.cuh -- k_fse2: sequence decode with the serial chain and the value work on two warps (included by kernels.cu).
//
// decode_sequences (sequence_section_decoder.rs:14-221) is one dependency chain per block: the three FSE states and the
// bit position.  Everything else per sequence -- code -> (baseline, extra bits), extracting the extra bits, the values,
// do_offset_history, the prefix sums, the record -- hangs off the chain but is not part of it.  Round 1's k_fse did
// both in one instruction stream (106 warp-instructions per sequence, the longest block's chain = the kernel's time).
// Here a CTA is two warps, one lane per block in each:
//   * the CHAIN warp walks states and bit position only (window, the two extra-bit look-ups, three table look-ups) and
//     drops {LL entry, ML entry, OF entry, the 32 window bits holding the extra bits} per sequence into a shared-memory
//     queue, one slot row per step, lane-interleaved (conflict-free);
//   * the VALUE warp trails it through the queue and computes ll / ml / offset, runs do_offset_history on symbolic history
//     values, keeps the prefix sums and writes the records (four sequences = three 16-byte stores).
// The two warps sit on different schedulers of the SM, so the chain's step shrinks to its own instructions.
// Anything unusual (under/over-run, > 32 extra bits in one sequence, offset code >= 30, bad table) is noticed by the chain
// warp; such a block is decoded again by fse_exact_block (the reference's control flow), exactly as in k_fse.
#pragma once

namespace b200z {

constexpr uint32_t F2_LANES = 32;                                   // blocks per CTA (one lane of each warp per block)
constexpr uint32_t F2_QDEPTH = 32;                                  // queue depth in steps (multiple of 4)
constexpr uint32_t F2_OFF_LUT = F2_LANES * FSE_TAB_U16 * 2;
constexpr uint32_t F2_OFF_RING = F2_OFF_LUT + 1024;
constexpr uint32_t F2_OFF_QUEUE = F2_OFF_RING + F2_LANES * RING_STRIDE;
constexpr uint32_t F2_OFF_CTRL = F2_OFF_QUEUE + F2_QDEPTH * 3 * 32 * 4;
constexpr uint32_t kFse2Smem = F2_OFF_CTRL + 64;
constexpr uint32_t F2_SPIN_LIMIT = 1u << 24;

struct F2Ctrl { uint32_t prod, cons, bad_mask, final_ready, abort; };

__device__ __forceinline__ uint32_t f2_ld_acquire(uint32_t a) { uint32_t v; asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void f2_st_release(uint32_t a, uint32_t v) { asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__global__ void __launch_bounds__(64) k_fse2(const BlockDesc *__restrict__ descs, BlockAux *__restrict__ aux, const uint8_t *__restrict__ input,
                                            uint32_t *__restrict__ seq_scratch, uint32_t nblocks) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // this CTA is resident: k_exec may follow
    extern __shared__ __align__(16) uint8_t smem_f2[];
    uint16_t *tabs = reinterpret_cast<uint16_t *>(smem_f2);
    uint32_t *s_ll_base = reinterpret_cast<uint32_t *>(smem_f2 + F2_OFF_LUT);
    uint32_t *s_ml_base = s_ll_base + 36;
    uint8_t *s_ll_bits = reinterpret_cast<uint8_t *>(s_ml_base + 53);
    uint8_t *s_ml_bits = s_ll_bits + 36;
    uint32_t *s_ll = reinterpret_cast<uint32_t *>(smem_f2 + F2_OFF_LUT + 512);   // base | bits << 24
    uint32_t *s_ml = s_ll + 36;
    const uint32_t S = (uint32_t)__cvta_generic_to_shared(smem_f2);
    const uint32_t S_q = S + F2_OFF_QUEUE, S_ctrl = S + F2_OFF_CTRL;
    const uint32_t S_prod = S_ctrl + 0, S_cons = S_ctrl + 4, S_bad = S_ctrl + 8, S_final = S_ctrl + 12, S_abort = S_ctrl + 16;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint32_t i = tid; i < 36; i += 64) { s_ll_base[i] = c_ll_base[i]; s_ll_bits[i] = c_ll_bits[i]; s_ll[i] = c_ll_base[i] | ((uint32_t)c_ll_bits[i] << 24); }
    for (uint32_t i = tid; i < 53; i += 64) { s_ml_base[i] = c_ml_base[i]; s_ml_bits[i] = c_ml_bits[i]; s_ml[i] = c_ml_base[i] | ((uint32_t)c_ml_bits[i] << 24); }
    if (tid < 8) sts32(S_ctrl + (tid << 2), 0u);

    // ---- my block (both warps look at the same block per lane)
    const uint32_t b = blockIdx.x * F2_LANES + lane;
    const bool active = b < nblocks;
    const BlockDesc *d = active ? &descs[b] : nullptr;
    uint32_t st_seq = 0;
    bool run = false;
    if (active) {
        st_seq = aux[b].pad;
        if (warp == 0) {
            if (d->btype != BT_COMPRESSED) aux[b].out_size = d->raw_size;
            else if (!d->host_status && d->nseq == 0) aux[b].out_size = d->regen_size;
        }
        run = d->btype == BT_COMPRESSED && !d->host_status && d->nseq != 0 && st_seq == 0;
    }
    const FseTab *tl = run ? d->ll : nullptr, *to = run ? d->of : nullptr, *tm = run ? d->ml : nullptr;
    bool bad = run && (!tl || !tl->valid || !to || !to->valid || !tm || !tm->valid);
    if (run && !bad) bad = tl->log > 9u || tm->log > 9u || to->log > 8u;
    // ---- stage the tables of the CTA's blocks (warp w takes every second block; 16-byte vectors)
    for (uint32_t j = warp; j < F2_LANES; j += 2) {
        const FseTab *pj[3];
        pj[0] = (const FseTab *)(uintptr_t)__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)(bad ? nullptr : tl), j);
        pj[1] = (const FseTab *)(uintptr_t)__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)(bad ? nullptr : tm), j);
        pj[2] = (const FseTab *)(uintptr_t)__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)(bad ? nullptr : to), j);
        uint16_t *dstj = tabs + j * FSE_TAB_U16;
        const uint32_t offs[3] = {0, 512, 1024};
#pragma unroll
        for (int t = 0; t < 3; t++) {
            if (!pj[t]) continue;
            uint32_t n16 = ((2u << pj[t]->log) + 15) >> 4;   // bytes / 16
            if (t == 2 && n16 > 32) n16 = 32;
            if (n16 > 64) n16 = 64;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(pj[t]->e);
            uint4 *d4 = reinterpret_cast<uint4 *>(dstj + offs[t]);
            for (uint32_t i = lane; i < n16; i += 32) d4[i] = s4[i];
        }
    }
    __syncthreads();
    const uint32_t qLL = (uint32_t)__cvta_generic_to_shared(s_ll), qML = (uint32_t)__cvta_generic_to_shared(s_ml);
    const uint32_t nseq = run ? d->nseq : 0u;
    const uint32_t aTL = S + lane * FSE_TAB_U16 * 2u, aTM = aTL + 1024u, aTO = aTL + 2048u;

    if (warp == 0) {
        // ============================================================ chain warp
        PosRing br;
        const uint32_t ring_addr = S + F2_OFF_RING + lane * RING_STRIDE + 16u;
        uint32_t logL = 0, logM = 0, logO = 0;
        // a lane without work walks a harmless fixed point: entry f = 1 with log 0 reads 0 bits and lands on itself
        uint32_t eL = 1, eM = 1, eO = 1;
        uint32_t qTL = aTL - 2u, qTM = aTM - 2u, qTO = aTO - 2u;
        bool alive = false;
        if (run && !bad) {
            const uint8_t *src = input + d->src_off + aux[b].seq_bits_off;
            const uint32_t len = d->src_size - aux[b].seq_bits_off;
            bad = !br.init(src, len, ring_addr);
            if (!bad) {
                logL = tl->log; logM = tm->log; logO = to->log;
                qTL = aTL - (2u << logL); qTM = aTM - (2u << logM); qTO = aTO - (2u << logO);
                // initial states LL, OF, ML (sequence_section_decoder.rs:164-166)
                uint32_t hi, lo;
                br.window(hi, lo);
                const uint32_t t1 = shl_c(hi, logL), t2 = shl_c(t1, logO);
                eL = fse_lds16(aTL + (shr_c(hi, 32u - logL) << 1));
                eO = fse_lds16(aTO + (shr_c(t1, 32u - logO) << 1));
                eM = fse_lds16(aTM + (shr_c(t2, 32u - logM) << 1));
                br.P -= (int32_t)(logL + logO + logM);
                alive = true;
            }
        }
        if (!alive) { br.ring = ring_addr; br.base = nullptr; br.gm1 = -1; br.next_g = -1; br.P = 0; }
        uint32_t steps = alive ? nseq : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));
        uint32_t flags = 0, max_of = 0, max_x = 0;
        bool aborted = false;
        for (uint32_t i = 0; i < steps && !aborted; i += 4) {
            if (i + 4 > F2_QDEPTH) {   // the slots of steps i .. i+3 are free once the value warp is done with step i + 3 - QDEPTH
                uint32_t spins = 0;
                while (f2_ld_acquire(S_cons) + F2_QDEPTH < i + 4) { if (++spins > F2_SPIN_LIMIT) { aborted = true; break; } }
                if (aborted) break;
            }
#pragma unroll
            for (uint32_t q = 0; q < 4; q++) {
                const uint32_t step = i + q;
                const bool act = alive && step < nseq, upd = act && step + 1 < nseq;
                uint32_t hi, lo;
                br.window(hi, lo);
                const uint32_t cL = eL >> 10, cM = eM >> 10, cO = eO >> 10;
                const uint32_t xL = lds32(qLL + (cL << 2)) >> 24, xM = lds32(qML + (cM << 2)) >> 24;
                const uint32_t xsum = cO + xM + xL;
                max_x = act ? max(max_x, xsum) : max_x;
                max_of = act ? max(max_of, cO) : max_of;
                {   // hand the sequence over: entries + the window bits that hold its extra bits
                    const uint32_t slot = S_q + ((step & (F2_QDEPTH - 1u)) * 96u + lane) * 4u;
                    sts32(slot, eL | (eM << 16)); sts32(slot + 128u, eO); sts32(slot + 256u, hi);
                }
                // state updates LL, ML, OF (:198-207); compact entries (b200z_types.h): nb = log - floor(log2 f)
                const uint32_t fL = eL & 1023u, fM = eM & 1023u, fO = eO & 1023u;
                const uint32_t nbL = logL - bfind32(fL), nbM = logM - bfind32(fM), nbO = logO - bfind32(fO);
                const uint32_t u0 = fsl_c(lo, hi, xsum);                 // the 32 bits below the extra bits
                const uint32_t u1 = shl_c(u0, nbL), u2 = shl_c(u1, nbM);
                const uint32_t aL = shr_c(u0, 32u - nbL), aM = shr_c(u1, 32u - nbM), aO = shr_c(u2, 32u - nbO);
                const uint32_t nL = fse_lds16(qTL + (((fL << nbL) + aL) << 1));
                const uint32_t nM = fse_lds16(qTM + (((fM << nbM) + aM) << 1));
                const uint32_t nO = fse_lds16(qTO + (((fO << nbO) + aO) << 1));
                eL = upd ? nL : eL; eM = upd ? nM : eM; eO = upd ? nO : eO;
                br.P -= act ? (int32_t)(xsum + (upd ? nbL + nbM + nbO : 0u)) : 0;
                if (q & 1) br.service();
            }
            flags |= (uint32_t)(br.P < 0) | (uint32_t)(max_x > 32u) | ((max_of + 2u) >> 5);
            if (flags) alive = false;   // the block goes to the exact path; its lane idles from here
            __syncwarp();
            if (lane == 0) f2_st_release(S_prod, i + 4);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (run && !bad) bad = flags != 0 || br.P != 0;
        const uint32_t bad_mask = __ballot_sync(0xffffffffu, bad || aborted);
        if (lane == 0) { sts32(S_bad, bad_mask); if (aborted) sts32(S_abort, 1u); f2_st_release(S_final, 1u); }
        // (verdicts are written by the value warp, whose own stores to the block's records come first in its program order)
    } else {
        // ============================================================ value warp
        const bool mine = run && !bad;   // (the chain warp may still give the block up: bad_mask at the end)
        uint32_t steps = mine ? nseq : 0u;
        // same step count as the chain warp: it depends on PosRing::init too, which only the chain warp runs -- take the
        // upper bound (blocks with tables) and stop when the chain warp says it is done
        uint32_t h0 = 1u << SEQ_SYM_SHIFT, h1 = 2u << SEQ_SYM_SHIFT, h2 = 3u << SEQ_SYM_SHIFT;   // "slot k at the block's start"
        uint32_t out_end = 0, lit_end = 0, ovf = 0;
        uint32_t *out = mine ? seq_scratch + d->seq_buf_off * 3 : nullptr;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));
        bool done = false;
        for (uint32_t i = 0; i < steps && !done; i += 4) {
            {   // wait for the chain warp to have published steps i .. i+3 (or to have finished: fewer steps than the bound)
                uint32_t spins = 0;
                for (;;) {
                    if (f2_ld_acquire(S_prod) >= i + 4) break;
                    if (f2_ld_acquire(S_final)) { done = f2_ld_acquire(S_prod) < i + 4; break; }
                    if (++spins > F2_SPIN_LIMIT) { done = true; break; }
                }
                if (done) break;
            }
            uint32_t stage[12];
#pragma unroll
            for (uint32_t q = 0; q < 4; q++) {
                const uint32_t step = i + q;
                const bool act = mine && step < nseq;
                const uint32_t slot = S_q + ((step & (F2_QDEPTH - 1u)) * 96u + lane) * 4u;
                const uint32_t w0 = lds32(slot), w1 = lds32(slot + 128u), hi = lds32(slot + 256u);
                const uint32_t cL = (w0 & 0xFFFFu) >> 10, cM = w0 >> 26, cO = (w1 & 0xFFFFu) >> 10;
                const uint32_t vL = lds32(qLL + (min(cL, 35u) << 2)), vM = lds32(qML + (min(cM, 52u) << 2));   // base | extra_bits << 24
                const uint32_t xL = vL >> 24, xM = vM >> 24, xO = cO;
                // extra bits: OF, ML, LL (get_bits_triple, sequence_section_decoder.rs:185)
                const uint32_t t1 = shl_c(hi, xO), t2 = shl_c(t1, xM);
                const uint32_t obits = shr_c(hi, 32u - xO), ml_add = shr_c(t1, 32u - xM), ll_add = shr_c(t2, 32u - xL);
                uint32_t offset = obits + (1u << (cO & 31u));
                const uint32_t ll = (vL & 0xFFFFFFu) + ll_add, ml = (vM & 0xFFFFFFu) + ml_add;
                {   // do_offset_history (sequence_execution.rs:59-118), branch-free, on symbolic history values
                    const bool rep = offset <= 3u;
                    const uint32_t r = offset - 1u + (ll == 0u ? 1u : 0u);   // 0..3 when rep
                    const uint32_t h0m1 = (h0 >> SEQ_SYM_SHIFT) ? h0 + 1u : h0 - (h0 != 0u ? 1u : 0u);   // saturating_sub (:74); symbols count the decrements
                    uint32_t cand = h0;
                    cand = r == 1u ? h1 : cand;
                    cand = r == 2u ? h2 : cand;
                    cand = r == 3u ? h0m1 : cand;
                    const uint32_t actual = rep ? cand : offset - 3u;
                    const bool keep2 = rep && r <= 1u, keep1 = rep && r == 0u;
                    if (act) { h2 = keep2 ? h2 : h1; h1 = keep1 ? h1 : h0; h0 = actual; }
                    offset = actual;
                }
                if (act) { lit_end += ll; out_end += ll + ml; ovf |= out_end | lit_end; }
                stage[3 * q] = out_end; stage[3 * q + 1] = lit_end; stage[3 * q + 2] = offset;
            }
            if (mine && i < nseq) {   // (the block's record array is padded to four sequences)
                uint4 *o4 = reinterpret_cast<uint4 *>(out + 3 * i);
                o4[0] = make_uint4(stage[0], stage[1], stage[2], stage[3]);
                o4[1] = make_uint4(stage[4], stage[5], stage[6], stage[7]);
                o4[2] = make_uint4(stage[8], stage[9], stage[10], stage[11]);
            }
            __syncwarp();
            if (lane == 0) f2_st_release(S_cons, i + 4);
        }
        {   // the chain warp's verdict
            uint32_t spins = 0;
            while (!f2_ld_acquire(S_final)) { if (++spins > F2_SPIN_LIMIT) break; }
        }
        const uint32_t bad_mask = lds32(S_bad);
        const bool aborted = lds32(S_abort) != 0u || !f2_ld_acquire(S_final);
        if (run && !aborted && ((bad_mask >> lane) & 1u)) {
            const uint16_t *TL = tabs + lane * FSE_TAB_U16;
            fse_exact_block(d, aux, b, input, seq_scratch, TL, TL + 512, TL + 1024, tl, to, tm, s_ll_base, s_ml_base, s_ll_bits, s_ml_bits, st_seq);
        } else if (active && !run) {
            aux[b].pad = st_seq;
        } else if (run && aborted) {
            aux[b].pad = mk_status(B200Z_ERR_CUDA, B200Z_STAGE_SEQUENCES);
        } else if (mine && !aborted) {
            aux[b].pad = 0;
            aux[b].hist_after[0] = h0; aux[b].hist_after[1] = h1; aux[b].hist_after[2] = h2;
            aux[b].sum_ll = lit_end;
            aux[b].flags = (ovf >> 31) ? AUX_WIDE : 0u;
            aux[b].out_size = (ovf >> 31) ? 0xffffffffu : out_end - lit_end + d->regen_size;   // sum of ml + regenerated literals
        }
        if (active) fse_publish_ready(aux, b);   // hand-off to k_exec
    }
}

}  // namespace b200z
