<?php
namespace Rindow\Math\Matrix;

use LogicException;

/**
 * MatrixOperator that knows the 'cuda' accelerator: MatrixOperator::laAccelerated (src/MatrixOperator.php:1566-1579)
 * only knows 'clblast', and the reference's own tests add accelerators by subclassing exactly like this
 * (tests/RindowTest/Math/Matrix/LinearAlgebraCLTest.php:29-37).
 *
 *     $mo = new MatrixOperatorCuda(service: new MatlibCuda());
 *     $la = $mo->laAccelerated('cuda');      // LinearAlgebraCuda over NDArrayCuda
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
}
