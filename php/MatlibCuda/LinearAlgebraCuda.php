<?php
namespace Rindow\Math\Matrix;

use InvalidArgumentException;
use LogicException;
use Interop\Polite\Math\Matrix\NDArray;
use Interop\Polite\Math\Matrix\LinearBuffer;
use Interop\Polite\Math\Matrix\DeviceBuffer;
use Interop\Polite\Math\Matrix\OpenCL;
use Rindow\Math\Matrix\Drivers\Service;
use Rindow\Math\Matrix\Drivers\MatlibCuda\CudaEventList;

/**
 * The accelerated LA object of the B200 driver: the counterpart of LinearAlgebraCL (src/LinearAlgebraCL.php) over
 * NDArrayCuda.  Constructed from a queue and the service (:75-93); accelerated() is true (:234-237); array() uploads
 * host arrays and passes device arrays through (:321-352); toNDArray() downloads (:354-360); alloc() creates device
 * arrays (:441-460); allocHost()/zerosHost() stay on the host (:793-814); blocking()/finish() (:145-152, :250-253);
 * Scalar-valued calls (sum, asum, nrm2, amax, amin, max, min, imax, imin, iamax, iamin, dot) are inherited as they are
 * and return PHP numbers: LinearAlgebra declares them `: float` / `: int`, and PHP's covariant return types do not let a
 * subclass hand back the rank-0 device array LinearAlgebraCL returns when scalarNumeric is off (:220-227, :2526-2529).
 * scalarNumeric() therefore reports true and refuses false; `$la->scalar($la->sum($x))`, the form the reference's
 * accelerated tests use (tests/RindowTest/Math/Matrix/LinearAlgebraTest.php), works unchanged.  The scalar comes back
 * through a blocking d2h copy of one element, so these calls are synchronising in either blocking() mode.
 *
 * Every operation is INHERITED from LinearAlgebra (src/LinearAlgebra.php): its methods only do shape algebra and one
 * positional call on $this->blas / $this->math with `$X->buffer(), $X->offset()`, and CudaBlas / CudaMath take
 * CudaDeviceBuffers exactly as they take the host-indexable CudaBuffers -- the OpenCL variants' trailing
 * ($events,$waitEvents) have no counterpart because the driver is stream-ordered.  The four host element reads of
 * LinearAlgebra (`$XX[$offX+$i]`, :488-511, :1645-1658) work through CudaDeviceBuffer::offsetGet (one small d2h).
 * Drop into src/ next to LinearAlgebraCL.php.  Executable mirror: rindow_math_matrix_b200/linear_algebra_cuda.py.
 */
class LinearAlgebraCuda extends LinearAlgebra
{
    protected object $queue;
    protected object $context;
    protected bool $blocking = false;

    public function __construct(object $queue, Service $service, ?int $defaultFloatType=null)
    {
        parent::__construct($service, $defaultFloatType, Service::LV_ADVANCED);
        $this->queue = $queue;
        $this->context = $queue->getContext();
    }

    public function getConfig() : string { return 'CudaB200'; }
    public function getContext() : object { return $this->context; }
    public function getQueue() : object { return $this->queue; }
    public function accelerated() : bool { return true; }
    public function fp64() : bool { return true; }
    /** @return array<string> */
    public function deviceTypes() : array { return ['GPU']; }
    public function finish() : void { $this->queue->finish(); }
    public function newEventList() : object { return new CudaEventList(); }

    public function blocking(?bool $switch=null) : bool
    {
        if($switch!==null) { $this->blocking = $switch; }
        return $this->blocking;
    }

    public function scalarNumeric(?bool $switch=null) : bool
    {
        if($switch===false) {
            throw new LogicException('LinearAlgebraCuda returns scalars as numbers (return types of LinearAlgebra).');
        }
        return true;
    }

    public function array(mixed $array, ?int $dtype=null, ?int $flags=null) : NDArray
    {
        if($array instanceof NDArray) {
            $buffer = $array->buffer();
            if($buffer instanceof DeviceBuffer) {
                return $array;
            }
            if(!($buffer instanceof LinearBuffer)) {
                throw new InvalidArgumentException('Unsuppored buffer type.');
            }
        } elseif(is_array($array) || is_numeric($array) || is_bool($array)) {
            $array = new NDArrayPhp($array, dtype:$dtype, service:$this->service);
        } else {
            throw new InvalidArgumentException('input value must be NDArray or array');
        }
        $flags = ($flags ?? OpenCL::CL_MEM_READ_WRITE) | OpenCL::CL_MEM_COPY_HOST_PTR;
        return new NDArrayCuda($this->queue, $array->buffer(), dtype:$array->dtype(), shape:$array->shape(),
            offset:$array->offset(), flags:$flags, service:$this->service);
    }

    public function toNDArray(NDArray $ndarray) : NDArray
    {
        return ($ndarray instanceof NDArrayCL) ? $ndarray->toNDArray() : $ndarray;
    }

    public function scalar(mixed $array) : mixed
    {
        return ($array instanceof NDArray) ? $array->toArray() : $array;
    }

    /** @param array<int> $shape */
    public function alloc(array $shape, ?int $dtype=null, ?int $flags=null) : NDArray
    {
        return new NDArrayCuda($this->queue, null, $dtype ?? $this->defaultFloatType, $shape, null, $flags,
            service:$this->service);
    }

    /** @param array<int> $shape */
    public function allocHost(array $shape, ?int $dtype=null) : NDArray
    {
        return new NDArrayPhp(null, dtype:$dtype ?? $this->defaultFloatType, shape:$shape, service:$this->service);
    }

    public function zerosHost(NDArray $X) : NDArray
    {
        $this->math->zeros($X->size(), $X->buffer(), $X->offset(), 1);
        return $X;
    }
}
