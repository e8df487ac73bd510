<?php
namespace Rindow\Math\Matrix;

use InvalidArgumentException;
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
 * scalar-valued calls return rank-0 device arrays unless scalarNumeric(true) (:220-227, :2526-2529).
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
    protected bool $scalarNumeric = false;

    /** scalar-valued methods wrapped by __call-free overrides below */
    protected const INDEX_CALLS = ['imax','imin','iamax','iamin'];

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
        if($switch!==null) { $this->scalarNumeric = $switch; }
        return $this->scalarNumeric;
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

    /** rank-0 device array holding a scalar result (LinearAlgebraCL returns $R unless scalarNumeric) */
    protected function scalarResult(string $name, mixed $value, ?NDArray $like) : mixed
    {
        if($this->blocking) { $this->finish(); }
        if($this->scalarNumeric || !is_numeric($value)) {
            return $value;
        }
        $dtype = in_array($name, self::INDEX_CALLS) ? NDArray::int32
               : (($like!==null && $this->isFloat($like)) ? $like->dtype() : $this->defaultFloatType);
        return $this->array(new NDArrayPhp($value, dtype:$dtype, service:$this->service));
    }

    public function sum(NDArray $X) : mixed   { return $this->scalarResult('sum',   parent::sum($X),   $X); }
    public function asum(NDArray $X) : mixed  { return $this->scalarResult('asum',  parent::asum($X),  $X); }
    public function nrm2(NDArray $X) : mixed  { return $this->scalarResult('nrm2',  parent::nrm2($X),  $X); }
    public function amax(NDArray $X) : mixed  { return $this->scalarResult('amax',  parent::amax($X),  $X); }
    public function amin(NDArray $X) : mixed  { return $this->scalarResult('amin',  parent::amin($X),  $X); }
    public function max(NDArray $X) : mixed   { return $this->scalarResult('max',   parent::max($X),   $X); }
    public function min(NDArray $X) : mixed   { return $this->scalarResult('min',   parent::min($X),   $X); }
    public function imax(NDArray $X) : mixed  { return $this->scalarResult('imax',  parent::imax($X),  $X); }
    public function imin(NDArray $X) : mixed  { return $this->scalarResult('imin',  parent::imin($X),  $X); }
    public function iamax(NDArray $X) : mixed { return $this->scalarResult('iamax', parent::iamax($X), $X); }
    public function iamin(NDArray $X) : mixed { return $this->scalarResult('iamin', parent::iamin($X), $X); }
    public function dot(NDArray $X, NDArray $Y, ?bool $conj=null) : mixed
    {
        return $this->scalarResult('dot', parent::dot($X, $Y, $conj), $X);
    }
}
