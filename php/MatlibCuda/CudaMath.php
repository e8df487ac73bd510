<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

use Interop\Polite\Math\Matrix\NDArray;
use Interop\Polite\Math\Matrix\Buffer;
use InvalidArgumentException;
use Rindow\Math\Matrix\Drivers\MatlibPHP\PhpMath;

/**
 * math() object of the B200 service.  The hot-path methods run on the device through the C ABI;
 * everything else is inherited from PhpMath and runs on the host mirror of the CudaBuffer.
 */
class CudaMath extends PhpMath
{
    protected function onDevice(Buffer ...$buffers) : bool
    {
        foreach ($buffers as $b) {
            if (!($b instanceof CudaBuffer)) { return false; }
        }
        $dt = $buffers[0]->dtype();
        return $dt == NDArray::float32 || $dt == NDArray::float64;
    }

    public function add(bool $trans,int $m,int $n,float $alpha,
        Buffer $X,int $offsetX,int $incX,Buffer $A,int $offsetA,int $ldA) : void
    {
        if (!$this->onDevice($X,$A)) { parent::add($trans,$m,$n,$alpha,$X,$offsetX,$incX,$A,$offsetA,$ldA); return; }
        $cols = (!$trans) ? $n : $m;
        if ($offsetX+($cols-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        if ($offsetA+($m-1)*$ldA+($n-1)>=count($A))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        CudaFFI::check(CudaFFI::lib()->rb200_add($A->abiDtype(),$trans?1:0,$m,$n,$alpha,
            $X->devicePtr(),$offsetX,$incX,$A->devicePtrRW(),$offsetA,$ldA));
    }

    public function multiply(bool $trans,int $m,int $n,
        Buffer $X,int $offsetX,int $incX,Buffer $A,int $offsetA,int $ldA) : void
    {
        if (!$this->onDevice($X,$A)) { parent::multiply($trans,$m,$n,$X,$offsetX,$incX,$A,$offsetA,$ldA); return; }
        $cols = (!$trans) ? $n : $m;
        if ($offsetX+($cols-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        if ($offsetA+($m-1)*$ldA+($n-1)>=count($A))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        CudaFFI::check(CudaFFI::lib()->rb200_multiply($A->abiDtype(),$trans?1:0,$m,$n,
            $X->devicePtr(),$offsetX,$incX,$A->devicePtrRW(),$offsetA,$ldA));
    }

    protected function checkReduce(int $m,int $n,int $k,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if ($offsetA+$m*$n*$k>count($A))
            throw new InvalidArgumentException('Matrix A specification too large for buffer.');
        if ($offsetB+$m*$k>count($B))
            throw new InvalidArgumentException('Matrix B specification too large for buffer.');
    }

    public function reduceSum(int $m,int $n,int $k,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if (!$this->onDevice($A,$B)) { parent::reduceSum($m,$n,$k,$A,$offsetA,$B,$offsetB); return; }
        $this->checkReduce($m,$n,$k,$A,$offsetA,$B,$offsetB);
        CudaFFI::check(CudaFFI::lib()->rb200_reduce_sum($A->abiDtype(),$m,$n,$k,
            $A->devicePtr(),$offsetA,$B->devicePtrRW(),$offsetB));
    }

    public function reduceMax(int $m,int $n,int $k,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if (!$this->onDevice($A,$B)) { parent::reduceMax($m,$n,$k,$A,$offsetA,$B,$offsetB); return; }
        $this->checkReduce($m,$n,$k,$A,$offsetA,$B,$offsetB);
        CudaFFI::check(CudaFFI::lib()->rb200_reduce_max($A->abiDtype(),$m,$n,$k,
            $A->devicePtr(),$offsetA,$B->devicePtrRW(),$offsetB));
    }

    public function reduceArgMax(int $m,int $n,int $k,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if (!$this->onDevice($A) || !($B instanceof CudaBuffer)) { parent::reduceArgMax($m,$n,$k,$A,$offsetA,$B,$offsetB); return; }
        $this->checkReduce($m,$n,$k,$A,$offsetA,$B,$offsetB);
        CudaFFI::check(CudaFFI::lib()->rb200_reduce_argmax($A->abiDtype(),$m,$n,$k,
            $A->devicePtr(),$offsetA,$B->abiDtype(),$B->devicePtrRW(),$offsetB));
    }

    public function softmax(int $m,int $n,Buffer $A,int $offsetA,int $ldA) : void
    {
        if (!$this->onDevice($A)) { parent::softmax($m,$n,$A,$offsetA,$ldA); return; }
        if ($offsetA+($m-1)*$ldA+($n-1)>=count($A))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        CudaFFI::check(CudaFFI::lib()->rb200_softmax($A->abiDtype(),$m,$n,$A->devicePtrRW(),$offsetA,$ldA));
    }

    public function im2col2d(bool $reverse,Buffer $images,int $images_offset,int $images_size,int $batches,
        int $im_h,int $im_w,int $channels,int $filter_h,int $filter_w,int $stride_h,int $stride_w,
        bool $padding,bool $channels_first,int $dilation_h,int $dilation_w,bool $cols_channels_first,
        Buffer $cols,int $cols_offset,int $cols_size) : void
    {
        if (!($images instanceof CudaBuffer) || !($cols instanceof CudaBuffer)) {
            parent::im2col2d($reverse,$images,$images_offset,$images_size,$batches,$im_h,$im_w,$channels,
                $filter_h,$filter_w,$stride_h,$stride_w,$padding,$channels_first,$dilation_h,$dilation_w,
                $cols_channels_first,$cols,$cols_offset,$cols_size);
            return;
        }
        // same checks and messages as PhpMath::im2col2d (PhpMath.php:2208-2246)
        $images_buf_size = $batches*$im_h*$im_w*$channels;
        if ($images_size!=$images_buf_size || count($images)-$images_offset<$images_buf_size) {
            throw new InvalidArgumentException('images buffer size is invalid');
        }
        $out_h = intdiv(($im_h-($filter_h-1)*$dilation_h-1),$stride_h)+1;
        $out_w = intdiv(($im_w-($filter_w-1)*$dilation_w-1),$stride_w)+1;
        if ($out_h<=0 || $out_w<=0) {
            throw new InvalidArgumentException('Invalid shape or parameters.');
        }
        $out_buf_size = $padding ? $batches*$im_h*$filter_h*$im_w*$filter_w*$channels
                                 : $batches*$out_h*$filter_h*$out_w*$filter_w*$channels;
        if ($cols_size!=$out_buf_size || count($cols)-$cols_offset!=$out_buf_size) {
            throw new InvalidArgumentException('output buffer size is invalid');
        }
        $im = $reverse ? $images->devicePtrRW() : $images->devicePtr();
        $co = $reverse ? $cols->devicePtr() : $cols->devicePtrRW();
        CudaFFI::check(CudaFFI::lib()->rb200_im2col2d($images->abiDtype(),$reverse?1:0,$im,$images_offset,$images_size,
            $batches,$im_h,$im_w,$channels,$filter_h,$filter_w,$stride_h,$stride_w,$padding?1:0,$channels_first?1:0,
            $dilation_h,$dilation_w,$cols_channels_first?1:0,$co,$cols_offset,$cols_size));
    }
    /** gatherb (PhpMath.php:1209-1260): batched gather / scatter / scatterAdd */
    public function gatherb(bool $reverse,bool $addMode,int $batches,int $m,int $n,int $k,int $len,int $numClass,
        Buffer $A,int $offsetA,Buffer $X,int $offsetX,Buffer $B,int $offsetB) : void
    {
        if (!$this->anyOnDevice($A,$X,$B) || $A->dtype()!=$B->dtype()) {
            parent::gatherb($reverse,$addMode,$batches,$m,$n,$k,$len,$numClass,$A,$offsetA,$X,$offsetX,$B,$offsetB); return;
        }
        if ($offsetX+$batches*$n*$k > count($X)) {
            throw new InvalidArgumentException('Matrix X specification too large for buffer.');
        }
        if ($offsetA+$batches*$m*$numClass*$k*$len > count($A)) {
            throw new InvalidArgumentException('Matrix A specification too large for buffer.');
        }
        if ($offsetB+$batches*$m*$n*$k*$len > count($B)) {
            throw new InvalidArgumentException('Matrix B specification too large for buffer.');
        }
        $a = $reverse ? $A->devicePtrRW() : $A->devicePtr();
        $b = $reverse ? $B->devicePtr() : $B->devicePtrRW();
        CudaFFI::check(CudaFFI::lib()->rb200_gatherb($A->abiDtype(),$X->abiDtype(),$reverse?1:0,$addMode?1:0,$batches,$m,$n,$k,
            $len,$numClass,$a,$offsetA,$X->devicePtr(),$offsetX,$b,$offsetB));
    }

    /** gathernd (PhpMath.php:1271-1325): gatherND / scatterND / scatterNDAdd; paramShape is read on the host */
    public function gathernd(bool $reverse,bool $addMode,int $m,int $n,int $k,int $indexDepth,Buffer $paramShape,
        Buffer $A,int $offsetA,Buffer $X,int $offsetX,Buffer $B,int $offsetB) : void
    {
        if (!$this->anyOnDevice($A,$X,$B) || $A->dtype()!=$B->dtype()) {
            parent::gathernd($reverse,$addMode,$m,$n,$k,$indexDepth,$paramShape,$A,$offsetA,$X,$offsetX,$B,$offsetB); return;
        }
        $shape = CudaFFI::lib()->new("int64_t[$indexDepth]");
        $paramSize = 1;
        for ($h=0;$h<$indexDepth;$h++) { $shape[$h] = $paramShape[$h]; $paramSize *= $paramShape[$h]; }
        if ($m*$n*$indexDepth > count($X)) {
            throw new InvalidArgumentException('Matrix X specification too large for buffer.');
        }
        if ($offsetA+$m*$paramSize*$k > count($A)) {
            throw new InvalidArgumentException('Matrix A specification too large for buffer.');
        }
        if ($offsetB+$m*$n*$k > count($B)) {
            throw new InvalidArgumentException('Matrix B specification too large for buffer.');
        }
        $a = $reverse ? $A->devicePtrRW() : $A->devicePtr();
        $b = $reverse ? $B->devicePtr() : $B->devicePtrRW();
        CudaFFI::check(CudaFFI::lib()->rb200_gathernd($A->abiDtype(),$X->abiDtype(),$reverse?1:0,$addMode?1:0,$m,$n,$k,
            $indexDepth,$shape,$a,$offsetA,$X->devicePtr(),$b,$offsetB));
    }

    /** einsum (PhpMath.php:3321-3357): the size / stride buffers are read on the host */
    public function einsum(Buffer $sizeOfIndices,Buffer $A,int $offsetA,Buffer $ldA,Buffer $B,int $offsetB,Buffer $ldB,
        Buffer $C,int $offsetC,int $ndimC) : void
    {
        $depth = count($sizeOfIndices);
        if (!$this->anyOnDevice($A,$B,$C) || !in_array($A->dtype(),[NDArray::float32,NDArray::float64]) || $depth>16) {
            parent::einsum($sizeOfIndices,$A,$offsetA,$ldA,$B,$offsetB,$ldB,$C,$offsetC,$ndimC); return;
        }
        $s = CudaFFI::lib()->new("int64_t[$depth]"); $la = CudaFFI::lib()->new("int64_t[$depth]"); $lb = CudaFFI::lib()->new("int64_t[$depth]");
        for ($h=0;$h<$depth;$h++) { $s[$h] = $sizeOfIndices[$h]; $la[$h] = $ldA[$h]; $lb[$h] = $ldB[$h]; }
        $this->einsumDevice($depth,$ndimC,$s,$la,$lb,$A,$offsetA,$B,$offsetB,$C,$offsetC);
    }

    /** einsum4p1 (PhpMath.php:3360-3440): four output axes, one summed axis; C is stored from element 0 as in the reference */
    public function einsum4p1(int $dim0,int $dim1,int $dim2,int $dim3,int $dim4,
        Buffer $A,int $offsetA,int $ldA0,int $ldA1,int $ldA2,int $ldA3,int $ldA4,
        Buffer $B,int $offsetB,int $ldB0,int $ldB1,int $ldB2,int $ldB3,int $ldB4,
        Buffer $C,int $offsetC) : void
    {
        if (!$this->anyOnDevice($A,$B,$C) || $A->dtype()!=$B->dtype() || $A->dtype()!=$C->dtype() || !in_array($A->dtype(),[NDArray::float32,NDArray::float64])) {
            parent::einsum4p1($dim0,$dim1,$dim2,$dim3,$dim4,$A,$offsetA,$ldA0,$ldA1,$ldA2,$ldA3,$ldA4,
                $B,$offsetB,$ldB0,$ldB1,$ldB2,$ldB3,$ldB4,$C,$offsetC); return;
        }
        if (count($C) <= $offsetC+$dim0*$dim1*$dim2*$dim3-1) {
            throw new InvalidArgumentException('Matrix specification too large for buffer C.');
        }
        $s = CudaFFI::lib()->new("int64_t[5]"); $la = CudaFFI::lib()->new("int64_t[5]"); $lb = CudaFFI::lib()->new("int64_t[5]");
        foreach ([$dim0,$dim1,$dim2,$dim3,$dim4] as $h=>$v) { $s[$h] = $v; }
        foreach ([$ldA0,$ldA1,$ldA2,$ldA3,$ldA4] as $h=>$v) { $la[$h] = $v; }
        foreach ([$ldB0,$ldB1,$ldB2,$ldB3,$ldB4] as $h=>$v) { $lb[$h] = $v; }
        $this->einsumDevice(5,4,$s,$la,$lb,$A,$offsetA,$B,$offsetB,$C,0);
    }

    private function einsumDevice(int $depth,int $ndimC,$s,$la,$lb,Buffer $A,int $offsetA,Buffer $B,int $offsetB,Buffer $C,int $offsetC) : void
    {
        $maxA = 0; $maxB = 0; $sizeC = 1;
        for ($h=0;$h<$depth;$h++) {
            $maxA += ($s[$h]-1)*$la[$h]; $maxB += ($s[$h]-1)*$lb[$h];
            if ($h<$ndimC) { $sizeC *= $s[$h]; }
        }
        if (count($A) <= $offsetA+$maxA) {
            throw new InvalidArgumentException('Matrix specification too large for buffer A.');
        }
        if (count($B) <= $offsetB+$maxB) {
            throw new InvalidArgumentException('Matrix specification too large for buffer B.');
        }
        if (count($C) < $offsetC+$sizeC) {
            throw new InvalidArgumentException('Matrix specification too large for buffer C.');
        }
        CudaFFI::check(CudaFFI::lib()->rb200_einsum($A->abiDtype(),$depth,$ndimC,$s,$A->devicePtr(),$offsetA,$la,
            $B->devicePtr(),$offsetB,$lb,$C->devicePtrRW(),$offsetC));
    }

    /** topk (PhpMath.php:933-1058): the reference's min-heap per row, one warp per row */
    public function topk(int $m,int $n,Buffer $input,int $offsetInput,int $k,bool $sorted,
        Buffer $values,int $offsetValues,Buffer $indices,int $offsetIndices) : void
    {
        if (!$this->anyOnDevice($input,$values,$indices) || $input->dtype()!=$values->dtype() || $k*4*12 > 200*1024) {
            parent::topk($m,$n,$input,$offsetInput,$k,$sorted,$values,$offsetValues,$indices,$offsetIndices); return;
        }
        // the reference's own checks, in its order (PhpMath.php:1035-1047)
        $this->assertShapeParameter("m", $m);
        $this->assertShapeParameter("n", $n);
        $this->assertShapeParameter("k", $k);
        $this->assertMatrixBufferSpec("input", $input, $m,$n, $offsetInput, $n);
        $this->assertMatrixBufferSpec("values", $values, $m,$k, $offsetValues, $k);
        $this->assertMatrixBufferSpec("indices", $indices, $m,$k, $offsetIndices, $k);
        if (!$this->isIntegerDtype($indices->dtype())) {
            throw new InvalidArgumentException("indices must be integers");
        }
        if ($k>$n) {
            return;
        }
        CudaFFI::check(CudaFFI::lib()->rb200_topk($input->abiDtype(),$indices->abiDtype(),$m,$n,$input->devicePtr(),$offsetInput,
            $k,$sorted?1:0,$values->devicePtrRW(),$offsetValues,$indices->devicePtrRW(),$offsetIndices));
    }

    /** imagecopy (PhpMath.php:2877-2957) */
    public function imagecopy(int $height,int $width,int $channels,Buffer $A,int $offsetA,Buffer $B,int $offsetB,
        bool $channelsFirst,int $heightShift,int $widthShift,bool $verticalFlip,bool $horizontalFlip,bool $rgbFlip) : void
    {
        if (!$this->anyOnDevice($A,$B) || $A->dtype()!=$B->dtype() || ($rgbFlip && $channels<3)) {
            parent::imagecopy($height,$width,$channels,$A,$offsetA,$B,$offsetB,$channelsFirst,$heightShift,$widthShift,
                $verticalFlip,$horizontalFlip,$rgbFlip); return;
        }
        if ($width*$height*$channels+$offsetA > count($A)) {
            throw new InvalidArgumentException('Matrix specification too large for bufferA');
        }
        if ($width*$height*$channels+$offsetB > count($B)) {
            throw new InvalidArgumentException('Matrix specification too large for bufferB');
        }
        CudaFFI::check(CudaFFI::lib()->rb200_imagecopy($A->abiDtype(),$height,$width,$channels,$A->devicePtr(),$offsetA,
            $B->devicePtrRW(),$offsetB,$channelsFirst?1:0,$heightShift,$widthShift,$verticalFlip?1:0,$horizontalFlip?1:0,$rgbFlip?1:0));
    }

    /** matrixcopy (PhpMath.php:2778-2809): B = alpha * A or its transpose -- the omatcopy kernel */
    public function matrixcopy(bool $trans,int $m,int $n,float $alpha,Buffer $A,int $offsetA,int $ldA,Buffer $B,int $offsetB,int $ldB) : void
    {
        if (!$this->anyOnDevice($A,$B) || $A->dtype()!=$B->dtype() || !in_array($A->dtype(),[NDArray::float32,NDArray::float64])) {
            parent::matrixcopy($trans,$m,$n,$alpha,$A,$offsetA,$ldA,$B,$offsetB,$ldB); return;
        }
        $rows = $trans ? $n : $m; $cols = $trans ? $m : $n;
        if ($m<1||$n<1) {
            throw new InvalidArgumentException('Argument m / n must be greater than 0.');
        }
        if ($ldA<$n || $offsetA+($m-1)*$ldA+$n > count($A)) {
            throw new InvalidArgumentException('Matrix specification too large for bufferA.');
        }
        if ($ldB<$cols || $offsetB+($rows-1)*$ldB+$cols > count($B)) {
            throw new InvalidArgumentException('Matrix specification too large for bufferB.');
        }
        CudaFFI::check(CudaFFI::lib()->rb200_omatcopy($A->abiDtype(),101,$trans?112:111,$m,$n,$alpha,
            $A->devicePtr(),$offsetA,$ldA,$B->devicePtrRW(),$offsetB,$ldB));
    }

    /** cumsumb (PhpMath.php:1861-1904): running sum along the middle axis of [m,n,k] */
    public function cumsumb(int $m,int $n,int $k,Buffer $A,int $offsetA,bool $exclusive,bool $reverse,Buffer $B,int $offsetB) : void
    {
        if (!$this->anyOnDevice($A,$B) || $A->dtype()!=$B->dtype()) { parent::cumsumb($m,$n,$k,$A,$offsetA,$exclusive,$reverse,$B,$offsetB); return; }
        if ($offsetA+$m*$n*$k > count($A)) {
            throw new InvalidArgumentException('Matrix specification too large for bufferA.');
        }
        if ($offsetB+$m*$n*$k > count($B)) {
            throw new InvalidArgumentException('Matrix specification too large for bufferB.');
        }
        CudaFFI::check(CudaFFI::lib()->rb200_cumsumb($A->abiDtype(),$m,$n,$k,$A->devicePtr(),$offsetA,
            $exclusive?1:0,$reverse?1:0,$B->devicePtrRW(),$offsetB));
    }

    /** cumsum (PhpMath.php:1822-1859): 1-D running sum with increments */
    public function cumsum(int $n,Buffer $X,int $offsetX,int $incX,bool $exclusive,bool $reverse,Buffer $Y,int $offsetY,int $incY) : void
    {
        if (!$this->anyOnDevice($X,$Y) || $X->dtype()!=$Y->dtype()) { parent::cumsum($n,$X,$offsetX,$incX,$exclusive,$reverse,$Y,$offsetY,$incY); return; }
        if ($n<1||$incX<1||$incY<1) {
            throw new InvalidArgumentException('Argument n / inc must be greater than 0.');
        }
        if ($offsetX+($n-1)*$incX >= count($X)) {
            throw new InvalidArgumentException('Vector specification too large for bufferX.');
        }
        if ($offsetY+($n-1)*$incY >= count($Y)) {
            throw new InvalidArgumentException('Vector specification too large for bufferY.');
        }
        CudaFFI::check(CudaFFI::lib()->rb200_cumsum($X->abiDtype(),$n,$X->devicePtr(),$offsetX,$incX,
            $exclusive?1:0,$reverse?1:0,$Y->devicePtrRW(),$offsetY,$incY));
    }

    /** searchsorted (PhpMath.php:1780-1820) */
    public function searchsorted(int $m,int $n,Buffer $A,int $offsetA,int $ldA,Buffer $X,int $offsetX,int $incX,
        bool $right,Buffer $Y,int $offsetY,int $incY) : void
    {
        if (!$this->anyOnDevice($A,$X,$Y) || $A->dtype()!=$X->dtype()) { parent::searchsorted($m,$n,$A,$offsetA,$ldA,$X,$offsetX,$incX,$right,$Y,$offsetY,$incY); return; }
        if ($m<1||$incX<1||$incY<1) {
            throw new InvalidArgumentException('Argument m / inc must be greater than 0.');
        }
        if ($offsetA+($m-1)*$ldA+$n > count($A)) {
            throw new InvalidArgumentException('Matrix specification too large for bufferA.');
        }
        if ($offsetX+($m-1)*$incX >= count($X)) {
            throw new InvalidArgumentException('Vector specification too large for bufferX.');
        }
        if ($offsetY+($m-1)*$incY >= count($Y)) {
            throw new InvalidArgumentException('Vector specification too large for bufferY.');
        }
        CudaFFI::check(CudaFFI::lib()->rb200_searchsorted($A->abiDtype(),$Y->abiDtype(),$m,$n,$A->devicePtr(),$offsetA,$ldA,
            $X->devicePtr(),$offsetX,$incX,$right?1:0,$Y->devicePtrRW(),$offsetY,$incY));
    }

    /** bandpart (PhpMath::bandpart): zero outside the band, any dtype */
    public function bandpart(int $m,int $n,int $k,Buffer $A,int $offset,int $lower,int $upper) : void
    {
        if (!$this->anyOnDevice($A)) { parent::bandpart($m,$n,$k,$A,$offset,$lower,$upper); return; }
        if ($offset+$m*$n*$k > count($A)) {
            throw new InvalidArgumentException('Matrix specification too large for buffer.');
        }
        CudaFFI::check(CudaFFI::lib()->rb200_bandpart($A->abiDtype(),$m,$n,$k,$A->devicePtrRW(),$offset,$lower,$upper));
    }

    /** masking (PhpMath.php:3229-3270) */
    public function masking(int $m,int $n,int $k,int $len,float $fill,int $mode,
        Buffer $X,int $offsetX,Buffer $A,int $offsetA) : void
    {
        if (!$this->anyOnDevice($X,$A)) {
            parent::masking($m,$n,$k,$len,$fill,$mode,$X,$offsetX,$A,$offsetA);
            return;
        }
        if ($X->dtype()!=NDArray::bool) {
            throw new InvalidArgumentException('dtype of X must be bool.');
        }
        if ($offsetX+$m*$k > count($X)) {
            throw new InvalidArgumentException('Matrix specification too large for buffer X.');
        }
        if ($offsetA+$m*$n*$k*$len > count($A)) {
            throw new InvalidArgumentException('Matrix specification too large for buffer A.');
        }
        CudaFFI::check(CudaFFI::lib()->rb200_masking($A->abiDtype(),$m,$n,$k,$len,$fill,$mode,
            $X->devicePtr(),$offsetX,$A->devicePtrRW(),$offsetA));
    }

    /** slice / stick (PhpMath.php:2691-2776): any dtype, bit-exact copies, += in add mode */
    public function slice(bool $reverse,bool $addMode,int $m,int $n,int $k,int $size,
        Buffer $A,int $offsetA,int $incA,Buffer $Y,int $offsetY,int $incY,
        int $startAxis0,int $sizeAxis0,int $startAxis1,int $sizeAxis1,int $startAxis2,int $sizeAxis2) : void
    {
        if (!$this->anyOnDevice($A,$Y) || $A->dtype()!=$Y->dtype()) {
            parent::slice($reverse,$addMode,$m,$n,$k,$size,$A,$offsetA,$incA,$Y,$offsetY,$incY,
                $startAxis0,$sizeAxis0,$startAxis1,$sizeAxis1,$startAxis2,$sizeAxis2);
            return;
        }
        if ($m*$n*$k*$size*$incA+$offsetA>count($A)) {
            throw new InvalidArgumentException('unmatch BufferA size and m,n,k');
        }
        if ($startAxis0<0||$startAxis0>=$m||$sizeAxis0<0||$sizeAxis0+$startAxis0>$m) {
            throw new InvalidArgumentException('Axis0 range is too large for source array.');
        }
        if ($startAxis1<0||$startAxis1>=$n||$sizeAxis1<0||$sizeAxis1+$startAxis1>$n) {
            throw new InvalidArgumentException('Axis1 range is too large for source array.');
        }
        if ($startAxis2<0||$startAxis2>=$k||$sizeAxis2<0||$sizeAxis2+$startAxis2>$k) {
            throw new InvalidArgumentException('Axis2 range is too large for source array.');
        }
        if ($sizeAxis0*$sizeAxis1*$sizeAxis2*$size*$incY>count($Y)-$offsetY) {
            throw new InvalidArgumentException('BufferY size is too small');
        }
        $a = $reverse ? $A->devicePtrRW() : $A->devicePtr();
        $y = $reverse ? $Y->devicePtr() : $Y->devicePtrRW();
        CudaFFI::check(CudaFFI::lib()->rb200_slice($A->abiDtype(),$reverse?1:0,$addMode?1:0,$m,$n,$k,$size,
            $a,$offsetA,$incA,$y,$offsetY,$incY,$startAxis0,$sizeAxis0,$startAxis1,$sizeAxis1,$startAxis2,$sizeAxis2));
    }

    /** repeat (PhpMath.php:1400-1422): B[m,repeats,k] = A[m,k] */
    public function repeat(int $m,int $k,int $repeats,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if (!$this->anyOnDevice($A,$B) || $A->dtype()!=$B->dtype()) {
            parent::repeat($m,$k,$repeats,$A,$offsetA,$B,$offsetB);
            return;
        }
        if ($offsetA+$m*$k>count($A)) {
            throw new InvalidArgumentException('Matrix A specification too large for buffer.');
        }
        if ($offsetB+$m*$repeats*$k>count($B)) {
            throw new InvalidArgumentException('Matrix B specification too large for buffer.');
        }
        CudaFFI::check(CudaFFI::lib()->rb200_repeat($A->abiDtype(),$m,$k,$repeats,$A->devicePtr(),$offsetA,$B->devicePtrRW(),$offsetB));
    }

    /** im2col1d (PhpMath.php:1967-2088) is im2col2d with one row of height 1: same layouts, same copy */
    public function im2col1d(bool $reverse,Buffer $images,int $images_offset,int $images_size,int $batches,
        int $im_w,int $channels,int $filter_w,int $stride_w,bool $padding,bool $channels_first,int $dilation_w,
        bool $cols_channels_first,Buffer $cols,int $cols_offset,int $cols_size) : void
    {
        if (!($images instanceof CudaBuffer) || !($cols instanceof CudaBuffer)) {
            parent::im2col1d($reverse,$images,$images_offset,$images_size,$batches,$im_w,$channels,$filter_w,$stride_w,
                $padding,$channels_first,$dilation_w,$cols_channels_first,$cols,$cols_offset,$cols_size);
            return;
        }
        $images_buf_size = $batches*$im_w*$channels;
        if ($images_size!=$images_buf_size || count($images)-$images_offset<$images_buf_size) {
            throw new InvalidArgumentException('images buffer size is invalid');
        }
        $out_w = intdiv(($im_w-($filter_w-1)*$dilation_w-1),$stride_w)+1;
        if ($out_w<=0) {
            throw new InvalidArgumentException('Invalid shape or parameters.');
        }
        $out_buf_size = $batches*($padding ? $im_w : $out_w)*$filter_w*$channels;
        if ($cols_size!=$out_buf_size || count($cols)-$cols_offset!=$out_buf_size) {
            throw new InvalidArgumentException('output buffer size is invalid');
        }
        $this->im2col2d($reverse,$images,$images_offset,$images_size,$batches,1,$im_w,$channels,1,$filter_w,1,$stride_w,
            $padding,$channels_first,1,$dilation_w,$cols_channels_first,$cols,$cols_offset,$cols_size);
    }

    /** im2col3d (PhpMath.php:2396-2600) */
    public function im2col3d(bool $reverse,Buffer $images,int $images_offset,int $images_size,int $batches,
        int $im_d,int $im_h,int $im_w,int $channels,int $filter_d,int $filter_h,int $filter_w,
        int $stride_d,int $stride_h,int $stride_w,bool $padding,bool $channels_first,
        int $dilation_d,int $dilation_h,int $dilation_w,bool $cols_channels_first,
        Buffer $cols,int $cols_offset,int $cols_size) : void
    {
        if (!($images instanceof CudaBuffer) || !($cols instanceof CudaBuffer)) {
            parent::im2col3d($reverse,$images,$images_offset,$images_size,$batches,$im_d,$im_h,$im_w,$channels,
                $filter_d,$filter_h,$filter_w,$stride_d,$stride_h,$stride_w,$padding,$channels_first,
                $dilation_d,$dilation_h,$dilation_w,$cols_channels_first,$cols,$cols_offset,$cols_size);
            return;
        }
        $images_buf_size = $batches*$im_d*$im_h*$im_w*$channels;
        if ($images_size!=$images_buf_size || count($images)-$images_offset<$images_buf_size) {
            throw new InvalidArgumentException('images buffer size is invalid');
        }
        $out_d = intdiv(($im_d-($filter_d-1)*$dilation_d-1),$stride_d)+1;
        $out_h = intdiv(($im_h-($filter_h-1)*$dilation_h-1),$stride_h)+1;
        $out_w = intdiv(($im_w-($filter_w-1)*$dilation_w-1),$stride_w)+1;
        if ($out_h<=0 || $out_w<=0) {
            throw new InvalidArgumentException('Invalid shape or parameters.');
        }
        $out_buf_size = $padding ? $batches*$im_d*$filter_d*$im_h*$filter_h*$im_w*$filter_w*$channels
                                 : $batches*$out_d*$filter_d*$out_h*$filter_h*$out_w*$filter_w*$channels;
        if ($cols_size!=$out_buf_size || count($cols)-$cols_offset!=$out_buf_size) {
            throw new InvalidArgumentException('output buffer size is invalid');
        }
        $im = $reverse ? $images->devicePtrRW() : $images->devicePtr();
        $co = $reverse ? $cols->devicePtr() : $cols->devicePtrRW();
        CudaFFI::check(CudaFFI::lib()->rb200_im2col3d($images->abiDtype(),$reverse?1:0,$im,$images_offset,$images_size,
            $batches,$im_d,$im_h,$im_w,$channels,$filter_d,$filter_h,$filter_w,$stride_d,$stride_h,$stride_w,
            $padding?1:0,$channels_first?1:0,$dilation_d,$dilation_h,$dilation_w,$cols_channels_first?1:0,
            $co,$cols_offset,$cols_size));
    }

    // ---- SURVEY.md section 8(f) widening: the rest of the broadcast family, the unary family, equal / notEqual /
    //      not, astype, sum / imax / imin, updateAddOnehot.  Op codes: RB200_BC_* / RB200_UN_* / RB200_CMP_*
    //      of include/rindow_b200.h.  Validation and messages are PhpMath's own (PhpMath.php:274-528, 687-722,
    //      860-891, 221-268, 611-682, 727-855, 1470-1547, 1710-1778, 150-216, 1438-1464).
    protected function anyOnDevice(Buffer ...$buffers) : bool
    {
        foreach ($buffers as $b) {
            if (!($b instanceof CudaBuffer)) { return false; }
        }
        return true;
    }

    protected function broadcastMN(string $name,int $op,int $m,int $n,
        Buffer $A,int $offsetA,int $ldA,Buffer $X,int $offsetX,int $incX) : void
    {
        if (!$this->onDevice($A,$X) || $A->dtype()!=$X->dtype()) {
            parent::$name($m,$n,$A,$offsetA,$ldA,$X,$offsetX,$incX); return;
        }
        if ($m<=0 || $n<=0) throw new InvalidArgumentException("m and n must be greater than 0");
        if (($m-1)*$ldA+($n-1)+$offsetA>=count($A)) throw new InvalidArgumentException("Matrix A is too small");
        if (($n-1)*$incX+$offsetX>=count($X)) throw new InvalidArgumentException("Buffer X is too small");
        CudaFFI::check(CudaFFI::lib()->rb200_broadcast($A->abiDtype(),$op,0,$m,$n,
            $A->devicePtrRW(),$offsetA,$ldA,$X->devicePtr(),$offsetX,$incX));
    }
    public function maximum(int $m,int $n,Buffer $A,int $offsetA,int $ldA,Buffer $X,int $offsetX,int $incX) : void
    { $this->broadcastMN('maximum',0,$m,$n,$A,$offsetA,$ldA,$X,$offsetX,$incX); }
    public function minimum(int $m,int $n,Buffer $A,int $offsetA,int $ldA,Buffer $X,int $offsetX,int $incX) : void
    { $this->broadcastMN('minimum',1,$m,$n,$A,$offsetA,$ldA,$X,$offsetX,$incX); }
    public function greater(int $m,int $n,Buffer $A,int $offsetA,int $ldA,Buffer $X,int $offsetX,int $incX) : void
    { $this->broadcastMN('greater',2,$m,$n,$A,$offsetA,$ldA,$X,$offsetX,$incX); }
    public function greaterEqual(int $m,int $n,Buffer $A,int $offsetA,int $ldA,Buffer $X,int $offsetX,int $incX) : void
    { $this->broadcastMN('greaterEqual',3,$m,$n,$A,$offsetA,$ldA,$X,$offsetX,$incX); }
    public function less(int $m,int $n,Buffer $A,int $offsetA,int $ldA,Buffer $X,int $offsetX,int $incX) : void
    { $this->broadcastMN('less',4,$m,$n,$A,$offsetA,$ldA,$X,$offsetX,$incX); }
    public function lessEqual(int $m,int $n,Buffer $A,int $offsetA,int $ldA,Buffer $X,int $offsetX,int $incX) : void
    { $this->broadcastMN('lessEqual',5,$m,$n,$A,$offsetA,$ldA,$X,$offsetX,$incX); }

    protected function broadcastTrans(int $op,bool $trans,int $m,int $n,
        Buffer $A,int $offsetA,int $ldA,Buffer $X,int $offsetX,int $incX) : void
    {
        $cols = (!$trans) ? $n : $m;
        if ($offsetX+($cols-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        if ($offsetA+($m-1)*$ldA+($n-1)>=count($A))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        CudaFFI::check(CudaFFI::lib()->rb200_broadcast($A->abiDtype(),$op,$trans?1:0,$m,$n,
            $A->devicePtrRW(),$offsetA,$ldA,$X->devicePtr(),$offsetX,$incX));
    }
    public function pow(bool $trans,int $m,int $n,Buffer $A,int $offsetA,int $ldA,Buffer $X,int $offsetX,int $incX) : void
    {
        if (!$this->onDevice($A,$X) || $A->dtype()!=$X->dtype()) { parent::pow($trans,$m,$n,$A,$offsetA,$ldA,$X,$offsetX,$incX); return; }
        $this->broadcastTrans(6,$trans,$m,$n,$A,$offsetA,$ldA,$X,$offsetX,$incX);
    }
    public function duplicate(bool $trans,int $m,int $n,Buffer $X,int $offsetX,int $incX,Buffer $A,int $offsetA,int $ldA) : void
    {
        if (!$this->onDevice($A,$X) || $A->dtype()!=$X->dtype()) { parent::duplicate($trans,$m,$n,$X,$offsetX,$incX,$A,$offsetA,$ldA); return; }
        $this->broadcastTrans(7,$trans,$m,$n,$A,$offsetA,$ldA,$X,$offsetX,$incX);
    }

    /** @return bool true when the device ran it */
    protected function unaryOnDevice(int $op,int $n,Buffer $X,int $offsetX,int $incX,float $alpha=1.0,float $beta=0.0) : bool
    {
        if (!$this->onDevice($X)) { return false; }
        if ($offsetX+($n-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        CudaFFI::check(CudaFFI::lib()->rb200_unary($X->abiDtype(),$op,$n,$alpha,$X->devicePtrRW(),$offsetX,$incX,$beta));
        return true;
    }
    public function increment(int $n,float $alpha,Buffer $X,int $offsetX,int $incX,float $beta) : void
    { if (!$this->unaryOnDevice(0,$n,$X,$offsetX,$incX,$alpha,$beta)) parent::increment($n,$alpha,$X,$offsetX,$incX,$beta); }
    public function reciprocal(int $n,float $alpha,Buffer $X,int $offsetX,int $incX,float $beta) : void
    { if (!$this->unaryOnDevice(1,$n,$X,$offsetX,$incX,$alpha,$beta)) parent::reciprocal($n,$alpha,$X,$offsetX,$incX,$beta); }
    public function square(int $n,Buffer $X,int $offsetX,int $incX) : void
    { if (!$this->unaryOnDevice(2,$n,$X,$offsetX,$incX)) parent::square($n,$X,$offsetX,$incX); }
    public function sqrt(int $n,Buffer $X,int $offsetX,int $incX) : void
    { if (!$this->unaryOnDevice(3,$n,$X,$offsetX,$incX)) parent::sqrt($n,$X,$offsetX,$incX); }
    public function rsqrt(int $n,float $alpha,Buffer $X,int $offsetX,int $incX,float $beta) : void
    { if (!$this->unaryOnDevice(4,$n,$X,$offsetX,$incX,$alpha,$beta)) parent::rsqrt($n,$alpha,$X,$offsetX,$incX,$beta); }
    public function exp(int $n,Buffer $X,int $offsetX,int $incX) : void
    { if (!$this->unaryOnDevice(5,$n,$X,$offsetX,$incX)) parent::exp($n,$X,$offsetX,$incX); }
    public function log(int $n,Buffer $X,int $offsetX,int $incX) : void
    { if (!$this->unaryOnDevice(6,$n,$X,$offsetX,$incX)) parent::log($n,$X,$offsetX,$incX); }
    public function tanh(int $n,Buffer $X,int $offsetX,int $incX) : void
    { if (!$this->unaryOnDevice(7,$n,$X,$offsetX,$incX)) parent::tanh($n,$X,$offsetX,$incX); }
    public function sin(int $n,Buffer $X,int $offsetX,int $incX) : void
    { if (!$this->unaryOnDevice(8,$n,$X,$offsetX,$incX)) parent::sin($n,$X,$offsetX,$incX); }
    public function cos(int $n,Buffer $X,int $offsetX,int $incX) : void
    { if (!$this->unaryOnDevice(9,$n,$X,$offsetX,$incX)) parent::cos($n,$X,$offsetX,$incX); }
    public function tan(int $n,Buffer $X,int $offsetX,int $incX) : void
    { if (!$this->unaryOnDevice(10,$n,$X,$offsetX,$incX)) parent::tan($n,$X,$offsetX,$incX); }
    public function nan2num(int $n,Buffer $X,int $offsetX,int $incX,float $alpha) : void
    { if (!$this->unaryOnDevice(12,$n,$X,$offsetX,$incX,$alpha)) parent::nan2num($n,$X,$offsetX,$incX,$alpha); }
    public function isnan(int $n,Buffer $X,int $offsetX,int $incX) : void
    { if (!$this->unaryOnDevice(13,$n,$X,$offsetX,$incX)) parent::isnan($n,$X,$offsetX,$incX); }
    public function not(int $n,Buffer $X,int $offsetX,int $incX) : void
    {
        if (!($X instanceof CudaBuffer)) { parent::not($n,$X,$offsetX,$incX); return; }
        if ($offsetX+($n-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        CudaFFI::check(CudaFFI::lib()->rb200_unary($X->abiDtype(),11,$n,1.0,$X->devicePtrRW(),$offsetX,$incX,0.0));
    }

    protected function compareOnDevice(int $op,int $n,Buffer $X,int $offsetX,int $incX,Buffer $Y,int $offsetY,int $incY) : bool
    {
        if (!$this->anyOnDevice($X,$Y) || $X->dtype()!=$Y->dtype()) { return false; }
        if ($offsetX+($n-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        if ($offsetY+($n-1)*$incY>=count($Y))
            throw new InvalidArgumentException('Vector specification too large for buffer.');
        CudaFFI::check(CudaFFI::lib()->rb200_compare($Y->abiDtype(),$op,$n,$X->devicePtr(),$offsetX,$incX,
            $Y->devicePtrRW(),$offsetY,$incY));
        return true;
    }
    public function equal(int $n,Buffer $X,int $offsetX,int $incX,Buffer $Y,int $offsetY,int $incY) : void
    { if (!$this->compareOnDevice(0,$n,$X,$offsetX,$incX,$Y,$offsetY,$incY)) parent::equal($n,$X,$offsetX,$incX,$Y,$offsetY,$incY); }
    public function notEqual(int $n,Buffer $X,int $offsetX,int $incX,Buffer $Y,int $offsetY,int $incY) : void
    { if (!$this->compareOnDevice(1,$n,$X,$offsetX,$incX,$Y,$offsetY,$incY)) parent::notEqual($n,$X,$offsetX,$incX,$Y,$offsetY,$incY); }

    public function astype(int $n,int $dtype,Buffer $X,int $offsetX,int $incX,Buffer $Y,int $offsetY,int $incY) : void
    {
        if (!$this->anyOnDevice($X,$Y)) { parent::astype($n,$dtype,$X,$offsetX,$incX,$Y,$offsetY,$incY); return; }
        CudaFFI::check(CudaFFI::lib()->rb200_astype($n,$X->abiDtype(),$X->devicePtr(),$offsetX,$incX,
            CudaBuffer::abiDtypeOf($dtype),$Y->devicePtrRW(),$offsetY,$incY));
    }

    public function sum(int $n,Buffer $X,int $offsetX,int $incX) : float
    {
        if (!($X instanceof CudaBuffer)) { return parent::sum($n,$X,$offsetX,$incX); }
        if ($offsetX+($n-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector X specification too large for buffer.');
        $out = CudaFFI::lib()->new('double');
        CudaFFI::check(CudaFFI::lib()->rb200_sum($X->abiDtype(),$n,$X->devicePtr(),$offsetX,$incX,\FFI::addr($out)));
        return $out->cdata;
    }
    public function imax(int $n,Buffer $X,int $offsetX,int $incX) : int
    {
        if (!($X instanceof CudaBuffer)) { return parent::imax($n,$X,$offsetX,$incX); }
        if ($offsetX+($n-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector X specification too large for buffer.');
        $out = CudaFFI::lib()->new('int64_t');
        CudaFFI::check(CudaFFI::lib()->rb200_imax($X->abiDtype(),$n,$X->devicePtr(),$offsetX,$incX,\FFI::addr($out)));
        return $out->cdata;
    }
    public function imin(int $n,Buffer $X,int $offsetX,int $incX) : int
    {
        if (!($X instanceof CudaBuffer)) { return parent::imin($n,$X,$offsetX,$incX); }
        if ($offsetX+($n-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector X specification too large for buffer.');
        $out = CudaFFI::lib()->new('int64_t');
        CudaFFI::check(CudaFFI::lib()->rb200_imin($X->abiDtype(),$n,$X->devicePtr(),$offsetX,$incX,\FFI::addr($out)));
        return $out->cdata;
    }

    public function updateAddOnehot(int $m,int $n,float $a,Buffer $X,int $offsetX,int $incX,Buffer $Y,int $offsetY,int $ldY) : void
    {
        if (!$this->anyOnDevice($X) || !$this->onDevice($Y)) { parent::updateAddOnehot($m,$n,$a,$X,$offsetX,$incX,$Y,$offsetY,$ldY); return; }
        if ($offsetX+($m-1)*$incX>=count($X))
            throw new InvalidArgumentException('Vector specification too large for bufferX.');
        if ($offsetY+($m-1)*$ldY+($n-1)>=count($Y))
            throw new InvalidArgumentException('Vector specification too large for bufferY.');
        $rc = CudaFFI::lib()->rb200_onehot($X->abiDtype(),$Y->abiDtype(),$m,$n,$a,
            $X->devicePtr(),$offsetX,$incX,$Y->devicePtrRW(),$offsetY,$ldY);
        if ($rc == -6) throw new \RuntimeException('Label number is out of bounds.');   // RB200_E_RANGE
        CudaFFI::check($rc);
    }
    // ---- gather / scatter / scatterAdd / reduceGather (PhpMath.php:1063-1208, 1362-1395) ------------------------
    public function gather(bool $reverse,bool $addMode,int $n,int $k,int $numClass,
        Buffer $X,int $offsetX,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if (!$this->anyOnDevice($X,$A,$B) || $A->dtype()!=$B->dtype()) {
            parent::gather($reverse,$addMode,$n,$k,$numClass,$X,$offsetX,$A,$offsetA,$B,$offsetB); return;
        }
        if ($offsetX+$n>count($X)) throw new InvalidArgumentException('Matrix X specification too large for buffer.');
        if ($offsetA+$numClass*$k>count($A)) throw new InvalidArgumentException('Matrix A specification too large for buffer.');
        if ($offsetB+$n*$k>count($B)) throw new InvalidArgumentException('Matrix B specification too large for buffer.');
        if ($numClass<=0) throw new InvalidArgumentException('numClass must be grator than zero.');
        CudaFFI::check(CudaFFI::lib()->rb200_gather($A->abiDtype(),$X->abiDtype(),$reverse?1:0,$addMode?1:0,$n,$k,$numClass,
            $X->devicePtr(),$offsetX,$reverse?$A->devicePtrRW():$A->devicePtr(),$offsetA,
            $reverse?$B->devicePtr():$B->devicePtrRW(),$offsetB));
    }

    public function reduceGather(bool $reverse,bool $addMode,int $m,int $n,int $numClass,
        Buffer $X,int $offsetX,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if (!$this->anyOnDevice($X,$A,$B) || $A->dtype()!=$B->dtype()) {
            parent::reduceGather($reverse,$addMode,$m,$n,$numClass,$X,$offsetX,$A,$offsetA,$B,$offsetB); return;
        }
        if ($offsetX+$m*$n>count($X)) throw new InvalidArgumentException('Matrix X specification too large for buffer.');
        if ($offsetA+$m*$numClass*$n>count($A)) throw new InvalidArgumentException('Matrix A specification too large for buffer.');
        if ($offsetB+$m*$n>count($B)) throw new InvalidArgumentException('Matrix B specification too large for buffer.');
        if ($numClass<=0) throw new InvalidArgumentException('numClass must be grator than zero.');
        CudaFFI::check(CudaFFI::lib()->rb200_reduce_gather($A->abiDtype(),$X->abiDtype(),$reverse?1:0,$addMode?1:0,$m,$n,$numClass,
            $X->devicePtr(),$offsetX,$reverse?$A->devicePtrRW():$A->devicePtr(),$offsetA,
            $reverse?$B->devicePtr():$B->devicePtrRW(),$offsetB));
    }

    // ---- N-D transpose (PhpMath.php:2959-3025): shape and perm are small int32 buffers, read on the host -----------
    public function transpose(Buffer $sourceShape,Buffer $perm,Buffer $A,int $offsetA,Buffer $B,int $offsetB) : void
    {
        if (!$this->anyOnDevice($A,$B)) { parent::transpose($sourceShape,$perm,$A,$offsetA,$B,$offsetB); return; }
        $ndim = count($sourceShape);
        if ($ndim!=count($perm)) throw new InvalidArgumentException('unmatch sourceshape and perm');
        if ($A->dtype()!=$B->dtype()) throw new InvalidArgumentException('unmatch sourceshape and perm');
        if ($ndim<=0) throw new InvalidArgumentException('Matrix must not be a scalar');
        $ffi = CudaFFI::lib();
        $shape = $ffi->new("int64_t[$ndim]"); $pm = $ffi->new("int64_t[$ndim]");
        $size = 1; $seen = [];
        for ($i=0;$i<$ndim;$i++) {
            $shape[$i] = $sourceShape[$i]; $pm[$i] = $perm[$i]; $size *= $sourceShape[$i];
            if ($perm[$i]>=$ndim) throw new InvalidArgumentException('dim value in the perm is out of range.');
            if (isset($seen[$perm[$i]])) throw new InvalidArgumentException('duplicate axis in perm option');
            $seen[$perm[$i]] = true;
        }
        if ($offsetA+$size>count($A)) throw new InvalidArgumentException('Matrix specification too large for bufferA');
        if ($offsetB+$size>count($B)) throw new InvalidArgumentException('Matrix specification too large for bufferB');
        CudaFFI::check($ffi->rb200_transpose($A->abiDtype(),$ndim,$shape,$pm,$A->devicePtr(),$offsetA,$B->devicePtrRW(),$offsetB));
    }

    // ---- strided-batched gemm: the entry LinearAlgebraCL::matmul calls on its math object (LinearAlgebraCL.php:2001);
    //      a LinearAlgebraCuda::matmul override issues it once per broadcast repeat instead of looping gemm ----------
    public function gemmStridedBatched(int $order,int $transA,int $transB,int $m,int $n,int $k,float $alpha,
        Buffer $A,int $offsetA,int $ldA,int $strideA,Buffer $B,int $offsetB,int $ldB,int $strideB,float $beta,
        Buffer $C,int $offsetC,int $ldC,int $strideC,int $batchSize) : void
    {
        if (!$this->onDevice($A,$B,$C)) throw new InvalidArgumentException('gemmStridedBatched needs device buffers');
        CudaFFI::check(CudaFFI::lib()->rb200_gemm_strided_batched($A->abiDtype(),$transA,$transB,$m,$n,$k,$alpha,
            $A->devicePtr(),$offsetA,$ldA,$strideA,$B->devicePtr(),$offsetB,$ldB,$strideB,$beta,
            $C->devicePtrRW(),$offsetC,$ldC,$strideC,$batchSize,0));
    }

    /**
     * Fused conv2d forward: out[b,oy,ox,f] = im2col2d(images) x kernel (+ bias) without a cols buffer
     * (rb200_conv2d_forward; the composition LinearAlgebra users write by hand is im2col :3369 -> gemm :925 -> add :1968,
     * and LinearAlgebraCL::im2col2dclblast :5001-5144 is the reference's own fused precedent).  Channels-last images
     * [b,h,w,c], kernel [kh,kw,c,filters].  Returns false -- nothing launched -- when the shape needs the two-call path
     * (status RB200_E_UNSUPPORTED): short reductions kh*kw*c <= 32 and hidden layers with channels % 32 == 0 are taken.
     */
    public function conv2dForward(Buffer $images,int $imagesOffset,int $batches,int $im_h,int $im_w,int $channels,
        int $filter_h,int $filter_w,int $stride_h,int $stride_w,bool $padding,int $dilation_h,int $dilation_w,
        Buffer $kernel,int $kernelOffset,int $filters,?Buffer $bias,int $biasOffset,Buffer $out,int $outOffset,
        int $gemmMode=0) : bool
    {
        if (!$this->onDevice($images) || !($kernel instanceof CudaBuffer) || !($out instanceof CudaBuffer)) { return false; }
        if (count($images)-$imagesOffset < $batches*$im_h*$im_w*$channels)
            throw new InvalidArgumentException('images buffer size is invalid');
        if (count($kernel)-$kernelOffset < $filter_h*$filter_w*$channels*$filters)
            throw new InvalidArgumentException('kernel buffer size is invalid');
        if ($bias!==null && count($bias)-$biasOffset < $filters)
            throw new InvalidArgumentException('bias buffer size is invalid');
        $rc = CudaFFI::lib()->rb200_conv2d_forward($images->abiDtype(),$images->devicePtr(),$imagesOffset,
            $batches,$im_h,$im_w,$channels,$filter_h,$filter_w,$stride_h,$stride_w,$padding?1:0,$dilation_h,$dilation_w,
            $kernel->devicePtr(),$kernelOffset,$filters,($bias instanceof CudaBuffer)?$bias->devicePtr():null,$biasOffset,
            $out->devicePtrRW(),$outOffset,$gemmMode);
        if ($rc == -3) { return false; }                  // RB200_E_UNSUPPORTED
        CudaFFI::check($rc);
        return true;
    }

    /**
     * (value, index) records of ONE STRIPE of a reduced axis that is sharded over processes, for the cross-rank merge
     * of math_imax (PhpMath.php:114-130): $records is an int64 CudaBuffer holding two words per output -- the bits of
     * the double value, then the stripe-local index.  $global0: the stripe starts at global position 0.
     */
    public function reduceArgMaxPartial(int $m,int $n,int $k,Buffer $A,int $offsetA,bool $global0,
        Buffer $records,int $offsetRecords) : void
    {
        if (!$this->onDevice($A) || !($records instanceof CudaBuffer))
            throw new InvalidArgumentException('reduceArgMaxPartial needs device buffers');
        if ($m<1||$n<1||$k<1) throw new InvalidArgumentException('m, n and k must be greater than 0.');
        if ($offsetA<0 || $offsetA+$m*$n*$k>count($A))
            throw new InvalidArgumentException('Matrix specification too large for bufferA.');
        if ($offsetRecords<0 || 2*($offsetRecords+$m*$k)>count($records))
            throw new InvalidArgumentException('records must be an int64 buffer of two words per output');
        CudaFFI::check(CudaFFI::lib()->rb200_reduce_argmax_partial($A->abiDtype(),$m,$n,$k,$A->devicePtr(),$offsetA,
            $global0?1:0,$records->devicePtrRW(),$offsetRecords));
    }
}
