<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

use Rindow\Math\Matrix\Drivers\AbstractMatlibService;

/**
 * The B200 service: an AbstractMatlibService with the three CUDA factories injected
 * (src/Drivers/AbstractMatlibService.php:43-75).  serviceLevel() then diagnoses LV_ADVANCED and
 * NDArrayPhp / LinearAlgebra / MatrixOperator use the CUDA driver with zero edits:
 *
 *     $mo = new MatrixOperator(service: new MatlibCuda());
 *     // or: new MatrixOperator(catalog: [MatlibCuda::class])      (src/Drivers/Selector.php:36-71)
 */
class MatlibCuda extends AbstractMatlibService
{
    protected string $name = 'matlib_cuda';

    protected function injectDefaultFactories() : void
    {
        $this->bufferFactory   ??= new CudaBufferFactory();
        $this->openblasFactory ??= new CudaBlasFactory();
        $this->mathFactory     ??= new CudaMathFactory();
    }
}
