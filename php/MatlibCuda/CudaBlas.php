<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

use Interop\Polite\Math\Matrix\BLAS;
use Interop\Polite\Math\Matrix\NDArray;
use Interop\Polite\Math\Matrix\Buffer;
use InvalidArgumentException;
use FFI;
use Rindow\Math\Matrix\Drivers\MatlibPHP\PhpBlas;
use Rindow\Math\Matrix\Drivers\MatlibPHP\Utils;

/**
 * blas() object of the B200 service.  gemm, scal, axpy, copy, swap, dot, asum, nrm2, iamax, iamin, gemv and omatcopy run on the device for
 * float32/float64 CudaBuffers; every other BLAS method (and every other dtype) is inherited from PhpBlas
 * and therefore runs on the host mirror of the CudaBuffer (ArrayAccess), exactly as SURVEY.md section
 * 8(b) prescribes for the rows marked "next".
 */
class CudaBlas extends PhpBlas
{
    use Utils;

    /** 0 auto (library default: 3xTF32), 1 FP32 SIMT, 2 TF32, 3 3xTF32 */
    protected int $gemmMode = 0;

    public function setGemmMode(int $mode) : void { $this->gemmMode = $mode; }
    public function getConfig() : string { return 'CudaBlas sm_100a'; }
    public function getCorename() : string { return 'B200'; }
    public function hasOmatcopy() : bool { return true; }

    public function gemm(
        int $order, int $transA, int $transB, int $m, int $n, int $k,
        float|object $alpha,
        Buffer $A, int $offsetA, int $ldA,
        Buffer $B, int $offsetB, int $ldB,
        float|object $beta,
        Buffer $C, int $offsetC, int $ldC ) : void
    {
        $dtype = $A->dtype();
        if (!($A instanceof CudaBuffer) || ($dtype != NDArray::float32 && $dtype != NDArray::float64)) {
            parent::gemm($order,$transA,$transB,$m,$n,$k,$alpha,$A,$offsetA,$ldA,$B,$offsetB,$ldB,$beta,$C,$offsetC,$ldC);
            return;
        }
        // same validation, same messages as PhpBlas::gemm (PhpBlas.php:861-881)
        if ($order == BLAS::ColMajor) {
            [$m,$n] = [$n,$m];
        } elseif ($order != BLAS::RowMajor) {
            throw new InvalidArgumentException('Invalid Order type');
        }
        [$tA,$conjA] = $this->codeToTrans($transA);
        [$tB,$conjB] = $this->codeToTrans($transB);
        $this->assertShapeParameter('m',$m);
        $this->assertShapeParameter('n',$n);
        $this->assertShapeParameter('k',$k);
        $this->assertMatrixBufferSpec("A", $A, (!$tA) ? $m : $k, (!$tA) ? $k : $m, $offsetA, $ldA);
        $this->assertMatrixBufferSpec("B", $B, (!$tB) ? $k : $n, (!$tB) ? $n : $k, $offsetB, $ldB);
        $this->assertMatrixBufferSpec("C", $C, $m, $n, $offsetC, $ldC);
        $alpha = $this->cleanFloatNumber($alpha,'alpha');
        $beta = $this->cleanFloatNumber($beta,'beta');
        $ffi = CudaFFI::lib();
        if ($dtype == NDArray::float32 && $this->gemmStreamed($transA,$transB,$m,$n,$k,$alpha,
                $A,$offsetA,$ldA,$B,$offsetB,$ldB,$beta,$C,$offsetC,$ldC,$tA,$tB)) {
            return;
        }
        $a = $A->devicePtr(); $b = $B->devicePtr(); $c = $C->devicePtrRW();
        if ($dtype == NDArray::float32) {
            CudaFFI::check($ffi->rb200_sgemm(BLAS::RowMajor,$transA,$transB,$m,$n,$k,$alpha,
                $ffi->cast('float*',$a),$offsetA,$ldA,$ffi->cast('float*',$b),$offsetB,$ldB,
                $beta,$ffi->cast('float*',$c),$offsetC,$ldC,$this->gemmMode));
        } else {
            CudaFFI::check($ffi->rb200_dgemm(BLAS::RowMajor,$transA,$transB,$m,$n,$k,$alpha,
                $ffi->cast('double*',$a),$offsetA,$ldA,$ffi->cast('double*',$b),$offsetB,$ldB,
                $beta,$ffi->cast('double*',$c),$offsetC,$ldC));
        }
    }

    /** null: bring C back eagerly iff the caller already indexes C from PHP (its mirror exists); true/false: always/never */
    protected ?bool $hostResultPrefetch = null;
    public function setHostResultPrefetch(?bool $flag) : void { $this->hostResultPrefetch = $flag; }

    /**
     * Host-operand form (rb200_sgemm_streamed): when an operand's current copy is in its pinned mirror and/or
     * the caller reads C from PHP, the transfers are pipelined with the multiply instead of
     * flush -> kernel -> fetch.  Returns false when the plain path should run.
     */
    protected function gemmStreamed(int $transA, int $transB, int $m, int $n, int $k, float $alpha,
        CudaBuffer $A, int $offsetA, int $ldA, CudaBuffer $B, int $offsetB, int $ldB, float $beta,
        CudaBuffer $C, int $offsetC, int $ldC, bool $tA, bool $tB) : bool
    {
        if ($m < 2048 || $A === $C || $B === $C || $A === $B) return false;
        $whole = fn(CudaBuffer $buf, int $off, int $rows, int $cols, int $ld) : bool =>
            $off == 0 && $ld == $cols && count($buf) == $rows * $cols;
        $upA = $A->hostPending() && $whole($A,$offsetA,(!$tA)?$m:$k,(!$tA)?$k:$m,$ldA);
        $upB = $B->hostPending() && $whole($B,$offsetB,(!$tB)?$k:$n,(!$tB)?$n:$k,$ldB);
        $wantC = $this->hostResultPrefetch ?? $C->hasMirror();
        $downC = $wantC && ($whole($C,$offsetC,$m,$n,$ldC) || $C->mirrorCurrent());
        if (!($upA || $upB || $downC)) return false;
        if ($A->hostPending() && !$upA) $A->devicePtr();           // plain flush
        if ($B->hostPending() && !$upB) $B->devicePtr();
        $cHostCurrent = 0;
        if ($C->hostPending()) {
            if ($whole($C,$offsetC,$m,$n,$ldC) && $downC) { $cHostCurrent = 1; $C->streamedUploaded(); }
            else $C->devicePtr();
        }
        if ($downC) $C->ensureMirror();
        [$dA,$hA] = $A->rawPointers(); [$dB,$hB] = $B->rawPointers(); [$dC,$hC] = $C->rawPointers();
        $ffi = CudaFFI::lib();
        $f = fn($p) => $p === null ? null : $ffi->cast('float*',$p);
        CudaFFI::check($ffi->rb200_sgemm_streamed(BLAS::RowMajor,$transA,$transB,$m,$n,$k,$alpha,
            $f($dA), $upA ? $f($hA) : null, $offsetA,$ldA, $f($dB), $upB ? $f($hB) : null, $offsetB,$ldB,
            $beta, $f($dC), $downC ? $f($hC) : null, $cHostCurrent, $offsetC,$ldC,$this->gemmMode));
        if ($upA) $A->streamedUploaded();
        if ($upB) $B->streamedUploaded();
        $C->streamedResult($downC);
        return true;
    }

    /** true when every buffer is a float32/float64 CudaBuffer, i.e. the call can stay on the device */
    protected function onDevice(Buffer ...$buffers) : bool
    {
        foreach ($buffers as $b) {
            if (!($b instanceof CudaBuffer)) return false;
            if ($b->dtype() != NDArray::float32 && $b->dtype() != NDArray::float64) return false;
        }
        return true;
    }

    public function scal(int $n, float|object $alpha, Buffer $X, int $offsetX, int $incX) : void
    {
        if (!$this->onDevice($X)) { parent::scal($n,$alpha,$X,$offsetX,$incX); return; }
        $this->assertShapeParameter('n',$n);
        $this->assertVectorBufferSpec('X', $X, $n, $offsetX, $incX);
        $alpha = $this->cleanFloatNumber($alpha,'alpha');
        CudaFFI::check(CudaFFI::lib()->rb200_scal($X->abiDtype(),$n,$alpha,$X->devicePtrRW(),$offsetX,$incX));
    }

    public function axpy(int $n, float|object $alpha,
        Buffer $X, int $offsetX, int $incX, Buffer $Y, int $offsetY, int $incY) : void
    {
        if (!$this->onDevice($X,$Y) || $X->dtype()!=$Y->dtype()) {
            parent::axpy($n,$alpha,$X,$offsetX,$incX,$Y,$offsetY,$incY); return;
        }
        $this->assertShapeParameter('n',$n);
        $this->assertVectorBufferSpec('X', $X, $n, $offsetX, $incX);
        $this->assertVectorBufferSpec('Y', $Y, $n, $offsetY, $incY);
        $alpha = $this->cleanFloatNumber($alpha,'alpha');
        CudaFFI::check(CudaFFI::lib()->rb200_axpy($X->abiDtype(),$n,$alpha,
            $X->devicePtr(),$offsetX,$incX,$Y->devicePtrRW(),$offsetY,$incY));
    }

    public function copy(int $n, Buffer $X, int $offsetX, int $incX, Buffer $Y, int $offsetY, int $incY) : void
    {
        if (!$this->onDevice($X,$Y) || $X->dtype()!=$Y->dtype()) {
            parent::copy($n,$X,$offsetX,$incX,$Y,$offsetY,$incY); return;
        }
        $this->assertShapeParameter('n',$n);
        $this->assertVectorBufferSpec('X', $X, $n, $offsetX, $incX);
        $this->assertVectorBufferSpec('Y', $Y, $n, $offsetY, $incY);
        CudaFFI::check(CudaFFI::lib()->rb200_copy($X->abiDtype(),$n,
            $X->devicePtr(),$offsetX,$incX,$Y->devicePtrRW(),$offsetY,$incY));
    }

    public function swap(int $n, Buffer $X, int $offsetX, int $incX, Buffer $Y, int $offsetY, int $incY) : void
    {
        if (!$this->onDevice($X,$Y) || $X->dtype()!=$Y->dtype()) {
            parent::swap($n,$X,$offsetX,$incX,$Y,$offsetY,$incY); return;
        }
        $this->assertShapeParameter('n',$n);
        $this->assertVectorBufferSpec('X', $X, $n, $offsetX, $incX);
        $this->assertVectorBufferSpec('Y', $Y, $n, $offsetY, $incY);
        CudaFFI::check(CudaFFI::lib()->rb200_swap($X->abiDtype(),$n,
            $X->devicePtrRW(),$offsetX,$incX,$Y->devicePtrRW(),$offsetY,$incY));
    }

    /** level-1 reductions on the device for every real dtype (PhpBlas.php:196-222, :266-348, :366-392) */
    protected function realDevice(Buffer ...$buffers) : bool
    {
        foreach ($buffers as $b) {
            if (!($b instanceof CudaBuffer) || $this->cistype($b->dtype())) return false;
        }
        return true;
    }

    protected function scalarResult(string $fn, string $ctype, int $n, Buffer $X, int $offsetX, int $incX) : float|int
    {
        $this->assertShapeParameter('n',$n);
        $this->assertVectorBufferSpec('X', $X, $n, $offsetX, $incX);
        $out = CudaFFI::lib()->new($ctype);
        CudaFFI::check(CudaFFI::lib()->$fn($X->abiDtype(),$n,$X->devicePtr(),$offsetX,$incX,\FFI::addr($out)));
        return $out->cdata;
    }

    public function asum(int $n, Buffer $X, int $offsetX, int $incX) : float
    {
        if (!$this->realDevice($X)) return parent::asum($n,$X,$offsetX,$incX);
        return $this->scalarResult('rb200_asum','double',$n,$X,$offsetX,$incX);
    }

    public function nrm2(int $n, Buffer $X, int $offsetX, int $incX) : float
    {
        if (!$this->realDevice($X)) return parent::nrm2($n,$X,$offsetX,$incX);
        return $this->scalarResult('rb200_nrm2','double',$n,$X,$offsetX,$incX);
    }

    public function iamax(int $n, Buffer $X, int $offsetX, int $incX) : int
    {
        if (!$this->realDevice($X)) return parent::iamax($n,$X,$offsetX,$incX);
        return $this->scalarResult('rb200_iamax','int64_t',$n,$X,$offsetX,$incX);
    }

    public function iamin(int $n, Buffer $X, int $offsetX, int $incX) : int
    {
        if (!$this->realDevice($X)) return parent::iamin($n,$X,$offsetX,$incX);
        return $this->scalarResult('rb200_iamin','int64_t',$n,$X,$offsetX,$incX);
    }

    public function dot(int $n, Buffer $X, int $offsetX, int $incX, Buffer $Y, int $offsetY, int $incY) : float|object
    {
        if (!$this->realDevice($X,$Y) || $X->dtype()!=$Y->dtype()) return parent::dot($n,$X,$offsetX,$incX,$Y,$offsetY,$incY);
        $this->assertShapeParameter('n',$n);
        $this->assertVectorBufferSpec('X', $X, $n, $offsetX, $incX);
        $this->assertVectorBufferSpec('Y', $Y, $n, $offsetY, $incY);
        $out = CudaFFI::lib()->new('double');
        CudaFFI::check(CudaFFI::lib()->rb200_dot($X->abiDtype(),$n,$X->devicePtr(),$offsetX,$incX,
            $Y->devicePtr(),$offsetY,$incY,\FFI::addr($out)));
        return $out->cdata;
    }

    public function gemv(int $order, int $trans, int $m, int $n, float|object $alpha,
        Buffer $A, int $offsetA, int $ldA, Buffer $X, int $offsetX, int $incX,
        float|object $beta, Buffer $Y, int $offsetY, int $incY ) : void
    {
        if (!$this->onDevice($A,$X,$Y) || $A->dtype()!=$X->dtype() || $A->dtype()!=$Y->dtype()) {
            parent::gemv($order,$trans,$m,$n,$alpha,$A,$offsetA,$ldA,$X,$offsetX,$incX,$beta,$Y,$offsetY,$incY);
            return;
        }
        // same validation order as PhpBlas::gemv (PhpBlas.php:773-788)
        if ($order == BLAS::ColMajor) {
            [$m,$n] = [$n,$m];
        } elseif ($order != BLAS::RowMajor) {
            throw new InvalidArgumentException('Invalid Order type');
        }
        [$t,$conj] = $this->codeToTrans($trans);
        $this->assertShapeParameter('m',$m);
        $this->assertShapeParameter('n',$n);
        $this->assertMatrixBufferSpec("A", $A, $m, $n, $offsetA, $ldA);
        $this->assertVectorBufferSpec('X', $X, (!$t) ? $n : $m, $offsetX, $incX);
        $this->assertVectorBufferSpec('Y', $Y, (!$t) ? $m : $n, $offsetY, $incY);
        $alpha = $this->cleanFloatNumber($alpha,'alpha');
        $beta = $this->cleanFloatNumber($beta,'beta');
        CudaFFI::check(CudaFFI::lib()->rb200_gemv($A->abiDtype(),BLAS::RowMajor,$trans,$m,$n,$alpha,
            $A->devicePtr(),$offsetA,$ldA,$X->devicePtr(),$offsetX,$incX,$beta,$Y->devicePtrRW(),$offsetY,$incY));
    }

    public function omatcopy(int $order, int $trans, int $m, int $n, float|object $alpha,
        Buffer $A, int $offsetA, int $ldA, Buffer $B, int $offsetB, int $ldB) : void
    {
        if (!$this->onDevice($A,$B)) {
            parent::omatcopy($order,$trans,$m,$n,$alpha,$A,$offsetA,$ldA,$B,$offsetB,$ldB); return;
        }
        // same validation order as PhpBlas::omatcopy (PhpBlas.php:1560-1577)
        if ($order == BLAS::ColMajor) {
            [$m,$n] = [$n,$m];
        } elseif ($order != BLAS::RowMajor) {
            throw new InvalidArgumentException('Invalid Order type');
        }
        [$t,$conj] = $this->codeToTrans($trans);
        $this->assertShapeParameter('m',$m);
        $this->assertShapeParameter('n',$n);
        $this->assertMatrixBufferSpec("A", $A, $m, $n, $offsetA, $ldA);
        $this->assertMatrixBufferSpec("B", $B, (!$t) ? $m : $n, (!$t) ? $n : $m, $offsetB, $ldB);
        if ($A->dtype()!=$B->dtype()) {
            throw new InvalidArgumentException("Unmatch data type for A and B");
        }
        $alpha = $this->cleanFloatNumber($alpha,'alpha');
        CudaFFI::check(CudaFFI::lib()->rb200_omatcopy($A->abiDtype(),BLAS::RowMajor,$trans,$m,$n,$alpha,
            $A->devicePtr(),$offsetA,$ldA,$B->devicePtrRW(),$offsetB,$ldB));
    }

    /**
     * Row-sharded gemm whose C stripe also lands in every other process's full C (one process per GPU).
     * $C[$offsetC...] is this rank's stripe inside its own full C; $peers are the other ranks' full-C allocations mapped
     * here (rb200_ipc_open of handles from rb200_ipc_export), or -- with $multicast, the NVLS multicast alias of a
     * symmetric allocation -- the single address a multimem store replicates through the NVSwitch.  Order the ranks
     * afterwards (a barrier of the process group).  Executable mirror: CudaBlas.gemmAllgather / sharding.RowShardedGemm.
     * @param array<int,mixed> $peers  void* device pointers
     */
    public function gemmAllgather(int $transA,int $transB,int $m,int $n,int $k,float $alpha,
        Buffer $A,int $offsetA,int $ldA,Buffer $B,int $offsetB,int $ldB,float $beta,
        Buffer $C,int $offsetC,int $ldC,array $peers,mixed $multicast=null) : void
    {
        $ffi = CudaFFI::lib();
        if ($multicast!==null) {
            CudaFFI::check($ffi->rb200_sgemm_allgather_mc(BLAS::RowMajor,$transA,$transB,$m,$n,$k,$alpha,
                $ffi->cast('float*',$A->devicePtr()),$offsetA,$ldA,$ffi->cast('float*',$B->devicePtr()),$offsetB,$ldB,
                $beta,$ffi->cast('float*',$C->devicePtrRW()),$offsetC,$ldC,$this->gemmMode,$multicast));
            return;
        }
        $np = count($peers);
        $arr = $ffi->new('void*['.max(1,$np).']');
        foreach (array_values($peers) as $i => $ptr) { $arr[$i] = $ptr; }
        CudaFFI::check($ffi->rb200_sgemm_allgather(BLAS::RowMajor,$transA,$transB,$m,$n,$k,$alpha,
            $ffi->cast('float*',$A->devicePtr()),$offsetA,$ldA,$ffi->cast('float*',$B->devicePtr()),$offsetB,$ldB,
            $beta,$ffi->cast('float*',$C->devicePtrRW()),$offsetC,$ldC,$this->gemmMode,$np,$arr));
    }

    /**
     * Record the driver calls issued by $calls into a CUDA graph and return a handle; graphLaunch() replays them with
     * one driver call (launch-bound sequences of small operations).  Issue the sequence once before capturing; no
     * allocating, scalar-returning or host-transfer calls inside.
     */
    public function graphCapture(callable $calls) : mixed
    {
        $ffi = CudaFFI::lib();
        CudaFFI::check($ffi->rb200_graph_begin());
        try { $calls(); } finally {
            $exec = $ffi->new('void*');
            $rc = $ffi->rb200_graph_end(\FFI::addr($exec));
        }
        CudaFFI::check($rc);
        return $exec;
    }
    public function graphLaunch(mixed $exec) : void { CudaFFI::check(CudaFFI::lib()->rb200_graph_launch($exec)); }
    public function graphDestroy(mixed $exec) : void { CudaFFI::check(CudaFFI::lib()->rb200_graph_destroy($exec)); }
}
