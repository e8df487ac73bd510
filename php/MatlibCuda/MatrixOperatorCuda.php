<?php
namespace Rindow\Math\Matrix;

use InvalidArgumentException;
use Interop\Polite\Math\Matrix\NDArray;

/**
 * MatrixOperator that knows the 'cuda' accelerator: MatrixOperator::laAccelerated (src/MatrixOperator.php:1566-1579)
 * only knows 'clblast', and the reference's own tests add accelerators by subclassing exactly like this
 * (tests/RindowTest/Math/Matrix/LinearAlgebraCLTest.php:29-37).
 *
 *     $mo = new MatrixOperatorCuda(service: new MatlibCuda());
 *     $la = $mo->laAccelerated('cuda');      // LinearAlgebraCuda over NDArrayCuda
 *
 * Reductions along an axis.  MatrixOperator::walkAxis (:824-857) asks the driver for ONE scalar per output element
 * (`$YY[$posY] = $math->sum($N,$XX,$offX+$bufPos,$incX)`): prod(shape)/shape[axis] driver calls.  Seen as A[m,n,k] with
 * n = shape[axis] that is the layout of the driver's reduce entry points (PhpMath.php:1552-1643), so for float arrays
 *   sum -> reduceSum;  mean -> reduceSum + scal(1/n);  argMax -> reduceArgMax (the math_imax rule, PhpMath.php:114-130);
 *   max -> reduceArgMax + reduceGather of X at those positions (`$XX[..+$pos*$incX]`, :903-912; not reduceMax, whose NaN
 *   rule is the sticky one);  argMin / min -> the same on -X (math_imin is math_imax mirrored)
 * run as one to three kernels.  Integer and bool arrays, the |x| family and axis=null keep the reference's walk.
 * Executable mirror with the reference's own expected values: rindow_math_matrix_b200/matrix_operator.py,
 * tests/test_matrix_operator_reductions.py.
 */
class MatrixOperatorCuda extends MatrixOperator
{
    protected ?object $cudaLA = null;

    /** @param array<string,mixed> $options */
    protected function createLinearAlgebraCuda(?array $options=null) : object
    {
        $queue = $this->service->createQueue($options);
        return new LinearAlgebraCuda($queue, $this->service);
    }

    /** @param array<string,mixed> $options */
    public function laAccelerated(string $name, ?array $options=null) : object
    {
        if($name==='cuda') {
            if($this->cudaLA===null) {
                $this->cudaLA = $this->createLinearAlgebraCuda($options);
            }
            return $this->cudaLA;
        }
        return parent::laAccelerated($name, $options);
    }

    protected function reducesOnDevice(NDArray $X, ?int $axis) : bool
    {
        if($axis===null || ($X->dtype()!=NDArray::float32 && $X->dtype()!=NDArray::float64)) {
            return false;
        }
        $shape = $X->shape();
        if($axis<0 || $axis>=count($shape)) {                       // same check, same message as walkAxis (:838-840)
            throw new InvalidArgumentException('Invalid axis: axis='.$axis.',shape=['.implode(',',$shape).']');
        }
        return true;
    }

    protected function argExtreme(NDArray $X, int $axis, bool $smallest, ?int $dtype=null) : NDArray
    {
        $la = $this->la();
        $src = $smallest ? $la->scal(-1.0, $la->copy($X)) : $X;
        return $la->reduceArgMax($src, axis:$axis, dtype:$dtype);
    }

    public function sum(NDArray $X, ?int $axis=null) : mixed
    {
        if(!$this->reducesOnDevice($X,$axis)) { return parent::sum($X,$axis); }
        return $this->la()->reduceSum($X, axis:$axis);
    }

    public function mean(NDArray $X, ?int $axis=null) : mixed
    {
        if(!$this->reducesOnDevice($X,$axis)) { return parent::mean($X,$axis); }
        $la = $this->la();
        return $la->scal(1.0/$X->shape()[$axis], $la->reduceSum($X, axis:$axis));
    }

    public function max(NDArray $X, ?int $axis=null) : mixed
    {
        if(!$this->reducesOnDevice($X,$axis)) { return parent::max($X,$axis); }
        return $this->la()->gather($X, $this->argExtreme($X,$axis,false), axis:$axis);
    }

    public function argMax(NDArray $X, ?int $axis=null) : mixed
    {
        if(!$this->reducesOnDevice($X,$axis)) { return parent::argMax($X,$axis); }
        return $this->argExtreme($X,$axis,false,$this->defaultIntType);
    }

    public function min(NDArray $X, ?int $axis=null) : mixed
    {
        if(!$this->reducesOnDevice($X,$axis)) { return parent::min($X,$axis); }
        return $this->la()->gather($X, $this->argExtreme($X,$axis,true), axis:$axis);
    }

    public function argMin(NDArray $X, ?int $axis=null) : mixed
    {
        if(!$this->reducesOnDevice($X,$axis)) { return parent::argMin($X,$axis); }
        return $this->argExtreme($X,$axis,true,$this->defaultIntType);
    }
}
