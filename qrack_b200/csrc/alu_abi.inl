// alu_abi.inl — C ABI of the QAlu family (included inside b200sv.cu's extern "C" block; declarations in include/b200sv.h).
// Argument checks and early returns follow src/qengine/arithmetic.cpp; measurement / SetReg pre-steps of the reference
// methods (MUL's SetReg(carry), IndexedADC's M(carry) ...) stay in the adapter above the ABI, as in the reference.

static bool alu_bad_range(const State* s, int start, int length)
{
    return start < 0 || length < 0 || start + length > s->nq;
}

// run one out-of-place basis map and adopt the new buffer
static int alu_run(State* s, AluDesc d, bool needZero, const unsigned char* hostTable, size_t tableBytes)
{
    SV_TRY(flush_queue(s));
    if (!s->amps) {
        return B200SV_OK; // zero state: nothing to permute (CHECK_ZERO_SKIP)
    }
    const uint64_t n = s->dim();
    const size_t bytes = (size_t)n * s->amp_bytes();
    // destination: the state's spare buffer (kept from the previous QAlu call) or a fresh allocation
    void* out = nullptr;
    cudaError_t e = cudaSuccess;
    if (s->spare && s->spare_bytes == bytes) {
        out = s->spare;
    } else {
        if (s->spare) {
            cudaStreamSynchronize(s->stream);
            cudaFree(s->spare);
        }
        s->spare = nullptr;
        s->spare_bytes = 0;
        e = cudaMalloc(&out, bytes);
        if (e != cudaSuccess) {
            return cuda_fail(e, "cudaMalloc(alu)");
        }
    }
    s->spare = out; // owned by the state from here on, whatever happens below
    s->spare_bytes = bytes;
    unsigned char* dtab = nullptr;
    if (hostTable && tableBytes) {
        e = cudaMalloc(&dtab, tableBytes);
        if (e == cudaSuccess) {
            e = cudaMemcpyAsync(dtab, hostTable, tableBytes, cudaMemcpyHostToDevice, s->stream);
        }
        if (e != cudaSuccess) {
            cudaFree(dtab);
            return cuda_fail(e, "alu table upload");
        }
    }
    d.table = dtab;
    if (needZero) {
        e = cudaMemsetAsync(out, 0, bytes, s->stream);
    }
    if (e == cudaSuccess) {
        const unsigned grid = stream_grid(s->dev, n, 256);
        if (s->prec == 32) {
            k_alu_map<float2><<<grid, 256, 0, s->stream>>>((const float2*)s->amps, (float2*)out, n, d);
        } else {
            k_alu_map<double2><<<grid, 256, 0, s->stream>>>((const double2*)s->amps, (double2*)out, n, d);
        }
        e = cudaGetLastError();
        s->stats.kernel_launches++;
        s->stats.bytes_swept += 2ULL * bytes;
    }
    if (e == cudaSuccess && s->external) {
        // the caller owns the buffer (P2P page / torch tensor): results go back into it
        e = cudaMemcpyAsync(s->amps, out, bytes, cudaMemcpyDeviceToDevice, s->stream);
    }
    if (dtab) {
        if (e == cudaSuccess) {
            e = cudaStreamSynchronize(s->stream); // the host table may go away as soon as we return
        }
        cudaFree(dtab);
    }
    if (e != cudaSuccess) {
        return cuda_fail(e, "alu map");
    }
    if (!s->external) {
        // ping-pong: the old buffer becomes the spare; everything is ordered on the state's stream, no sync needed
        s->spare = s->amps;
        s->amps = out;
    }
    return B200SV_OK;
}

static AluDesc alu_desc(int kind, int start, int length)
{
    AluDesc d;
    memset(&d, 0, sizeof(d));
    d.kind = kind;
    d.start = start;
    d.length = length;
    return d;
}

int b200sv_rol(b200sv_t s, int shift, int start, int length)
{
    SV_ENTER(s);
    if (alu_bad_range(s, start, length) || shift < 0) {
        return einval("ROL range is out-of-bounds!");
    }
    if (!length) {
        return B200SV_OK;
    }
    shift %= length;
    if (!shift) {
        return B200SV_OK;
    }
    // gather form: out[i] = in[rotr(i, shift)] — the strided side of the bit rotation is then the READ side (sector
    // over-fetch) instead of the WRITE side (partial-sector read-modify-write), ~2x faster on HBM
    AluDesc d = alu_desc(ALU_ROL, start, length);
    d.arg = (uint64_t)(length - shift);
    d.gather = 1;
    return alu_run(s, d, false, nullptr, 0);
}

int b200sv_inc(b200sv_t s, uint64_t to_add, int start, int length, uint64_t ctrl_mask)
{
    SV_ENTER(s);
    if (alu_bad_range(s, start, length)) {
        return einval("INC range is out-of-bounds!");
    }
    if (ctrl_mask >= s->dim()) {
        return einval("CINC control is out-of-bounds!");
    }
    if (!length) {
        return B200SV_OK;
    }
    const uint64_t lenMask = (1ULL << length) - 1U;
    to_add &= lenMask;
    if (!to_add) {
        return B200SV_OK;
    }
    AluDesc d = alu_desc(ALU_INC, start, length);
    d.arg = to_add;
    d.ctrlMask = ctrl_mask;
    return alu_run(s, d, false, nullptr, 0);
}

int b200sv_incdecc(b200sv_t s, uint64_t to_mod, int start, int length, int carry_index)
{
    SV_ENTER(s);
    if (alu_bad_range(s, start, length)) {
        return einval("INCDECC range is out-of-bounds!");
    }
    if (carry_index < 0 || carry_index >= s->nq) {
        return einval("INCDECC carryIndex is out-of-bounds!");
    }
    if (!length) {
        return B200SV_OK;
    }
    to_mod &= (1ULL << length) - 1U;
    if (!to_mod) {
        return B200SV_OK;
    }
    AluDesc d = alu_desc(ALU_INCC, start, length);
    d.arg = to_mod;
    d.carryMask = 1ULL << carry_index;
    d.zeroMask = d.carryMask;
    return alu_run(s, d, true, nullptr, 0);
}

int b200sv_incs(b200sv_t s, uint64_t to_add, int start, int length, int overflow_index)
{
    SV_ENTER(s);
    if (alu_bad_range(s, start, length)) {
        return einval("INCS range is out-of-bounds!");
    }
    if (overflow_index < 0 || overflow_index >= s->nq) {
        return einval("INCS overflowIndex is out-of-bounds!");
    }
    if (!length) {
        return B200SV_OK;
    }
    to_add &= (1ULL << length) - 1U;
    if (!to_add) {
        return B200SV_OK;
    }
    AluDesc d = alu_desc(ALU_INCS, start, length);
    d.arg = to_add;
    d.overflowMask = 1ULL << overflow_index;
    return alu_run(s, d, false, nullptr, 0);
}

int b200sv_incdecsc(b200sv_t s, uint64_t to_mod, int start, int length, int overflow_index, int carry_index)
{
    SV_ENTER(s);
    if (alu_bad_range(s, start, length)) {
        return einval("INCDECSC range is out-of-bounds!");
    }
    if (carry_index < 0 || carry_index >= s->nq) {
        return einval("INCDECSC carryIndex is out-of-bounds!");
    }
    if (overflow_index >= s->nq) {
        return einval("INCDECSC overflowIndex is out-of-bounds!");
    }
    if (!length) {
        return B200SV_OK;
    }
    to_mod &= (1ULL << length) - 1U;
    if (!to_mod) {
        return B200SV_OK;
    }
    AluDesc d = alu_desc(overflow_index < 0 ? ALU_INCSC : ALU_INCSC_OVF, start, length);
    d.arg = to_mod;
    d.carryMask = 1ULL << carry_index;
    d.zeroMask = d.carryMask;
    d.overflowMask = overflow_index < 0 ? 0 : (1ULL << overflow_index);
    return alu_run(s, d, true, nullptr, 0);
}

int b200sv_muldiv(b200sv_t s, int inverse, uint64_t to_mul, int start, int carry_start, int length, uint64_t ctrl_mask)
{
    SV_ENTER(s);
    if (alu_bad_range(s, start, length)) {
        return einval("MULDIV range is out-of-bounds!");
    }
    if (alu_bad_range(s, carry_start, length)) {
        return einval("MULDIV carry range is out-of-bounds!");
    }
    if (ctrl_mask >= s->dim()) {
        return einval("CMULDIV control is out-of-bounds!");
    }
    if (!length) {
        return B200SV_OK;
    }
    AluDesc d = alu_desc(inverse ? ALU_DIV : ALU_MUL, start, length);
    d.start2 = carry_start;
    d.length2 = length;
    d.arg = to_mul;
    d.ctrlMask = ctrl_mask;
    d.zeroMask = ((1ULL << length) - 1U) << carry_start;
    d.gather = inverse ? 1 : 0;
    return alu_run(s, d, true, nullptr, 0);
}

int b200sv_modnout(b200sv_t s, int kind, uint64_t to_mod, uint64_t mod_n, int in_start, int out_start, int length, uint64_t ctrl_mask)
{
    SV_ENTER(s);
    if (alu_bad_range(s, in_start, length)) {
        return einval("ModNOut inStart range is out-of-bounds!");
    }
    if (alu_bad_range(s, out_start, length)) {
        return einval("ModNOut outStart range is out-of-bounds!");
    }
    if (ctrl_mask >= s->dim()) {
        return einval("ModNOut control is out-of-bounds!");
    }
    if (kind < 0 || kind > 2 || !mod_n) {
        return einval("ModNOut: bad kind or modulus");
    }
    if (!length) {
        return B200SV_OK;
    }
    AluDesc d = alu_desc(kind == 0 ? ALU_MULMODN : (kind == 1 ? ALU_IMULMODN : ALU_POWMODN), in_start, length);
    d.start2 = out_start;
    d.length2 = length;
    d.arg = to_mod;
    d.modN = mod_n;
    d.ctrlMask = ctrl_mask;
    d.zeroMask = (((1ULL << length) - 1U) << out_start) & (s->dim() - 1U);
    d.gather = (kind == 1) ? 1 : 0;
    return alu_run(s, d, true, nullptr, 0);
}

int b200sv_indexed(b200sv_t s, int kind, int index_start, int index_length, int value_start, int value_length, int carry_index,
    int carry_in, const unsigned char* values)
{
    SV_ENTER(s);
    if (alu_bad_range(s, index_start, index_length)) {
        return einval("IndexedLDA/ADC/SBC index range is out-of-bounds!");
    }
    if (alu_bad_range(s, value_start, value_length)) {
        return einval("IndexedLDA/ADC/SBC value range is out-of-bounds!");
    }
    if (kind < 0 || kind > 2 || !values) {
        return einval("Indexed: bad kind or null table");
    }
    if (kind != 0 && (carry_index < 0 || carry_index >= s->nq)) {
        return einval("IndexedADC/SBC carryIndex is out-of-bounds!");
    }
    AluDesc d = alu_desc(kind == 0 ? ALU_LDA : (kind == 1 ? ALU_ADC : ALU_SBC), index_start, index_length);
    d.start2 = value_start;
    d.length2 = value_length;
    d.valueBytes = (value_length + 7) >> 3;
    d.carryIn = carry_in ? 1U : 0U;
    const uint64_t dimMask = s->dim() - 1U;
    if (kind == 0) {
        d.zeroMask = (((1ULL << value_length) - 1U) << value_start) & dimMask; // value register reads 0 (par_for_skip, :1078)
    } else {
        d.carryMask = 1ULL << carry_index;
        // ADC skips the carry qubit (:1251); SBC skips valueLength bits from the carry qubit upwards (:1436)
        d.zeroMask = (kind == 1) ? d.carryMask : ((((1ULL << value_length) - 1U) << carry_index) & dimMask);
    }
    const size_t tableBytes = ((size_t)1 << index_length) * (size_t)d.valueBytes;
    return alu_run(s, d, true, values, tableBytes);
}

int b200sv_hash(b200sv_t s, int start, int length, const unsigned char* values)
{
    SV_ENTER(s);
    if (alu_bad_range(s, start, length)) {
        return einval("Hash range is out-of-bounds!");
    }
    if (!values) {
        return einval("Hash: null table");
    }
    if (!length) {
        return B200SV_OK;
    }
    AluDesc d = alu_desc(ALU_HASH, start, length);
    d.valueBytes = (length + 7) >> 3;
    const size_t tableBytes = ((size_t)1 << length) * (size_t)d.valueBytes;
    return alu_run(s, d, true, values, tableBytes);
}

int b200sv_phase_flip_if_less(b200sv_t s, uint64_t greater_perm, int start, int length, int flag_index)
{
    SV_ENTER(s);
    if (alu_bad_range(s, start, length)) {
        return einval("PhaseFlipIfLess range is out-of-bounds!");
    }
    if (flag_index >= s->nq) {
        return einval("CPhaseFlipIfLess flagIndex is out-of-bounds!");
    }
    SV_TRY(flush_queue(s));
    if (!s->amps) {
        return B200SV_OK;
    }
    const uint64_t n = s->dim();
    const uint64_t regMask = ((1ULL << length) - 1U) << start;
    const uint64_t flagMask = flag_index < 0 ? 0 : (1ULL << flag_index);
    const unsigned grid = stream_grid(s->dev, n, 256);
    if (s->prec == 32) {
        k_phase_flip_if_less<float2><<<grid, 256, 0, s->stream>>>((float2*)s->amps, n, regMask, start, greater_perm, flagMask);
    } else {
        k_phase_flip_if_less<double2><<<grid, 256, 0, s->stream>>>((double2*)s->amps, n, regMask, start, greater_perm, flagMask);
    }
    SV_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return B200SV_OK;
}
