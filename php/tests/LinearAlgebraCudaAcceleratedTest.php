<?php
namespace RindowTest\Math\Matrix\LinearAlgebraCudaAcceleratedTest;

use Rindow\Math\Matrix\MatrixOperatorCuda;
use Rindow\Math\Matrix\NDArrayCL;
use Rindow\Math\Matrix\Drivers\Service;
use Rindow\Math\Matrix\Drivers\MatlibCuda\MatlibCuda;
// the reference's test classes are not autoloaded (cf. LinearAlgebraPHPModeTest.php:12-14)
if(!class_exists('RindowTest\Math\Matrix\LinearAlgebraTest\LinearAlgebraTest')) {
    require_once __DIR__.'/LinearAlgebraTest.php';
}
use RindowTest\Math\Matrix\LinearAlgebraTest\LinearAlgebraTest as ORGTest;

/**
 * The reference's LinearAlgebraTest over DEVICE-RESIDENT arrays, the way LinearAlgebraCLTest
 * (tests/RindowTest/Math/Matrix/LinearAlgebraCLTest.php:722-758) runs it over OpenCL: newLA() returns the accelerated
 * LA (blocking, numeric scalars), ndarray() downloads.  `$la->accelerated()` is true, so the "Large" cases the
 * reference skips on host drivers run (LinearAlgebraTest.php:3717, 3811, 7432, 7599, 7738, 7956, 9240) -- their
 * executable transcription is tests/test_gpu_device_ndarray.py of the B200 repository.
 */
class LinearAlgebraCudaAcceleratedTest extends ORGTest
{
    static protected $speedtest = false;

    public function setUp() : void
    {
        $this->service = new MatlibCuda();
        if($this->service->serviceLevel()<Service::LV_ACCELERATED) {
            $this->markTestSkipped('librindow_b200.so / B200 not available');
        }
    }

    public function newMatrixOperator()
    {
        return new MatrixOperatorCuda(service:$this->service);
    }

    public function newLA($mo)
    {
        $la = $mo->laAccelerated('cuda');
        $la->blocking(true);
        $la->scalarNumeric(true);
        return $la;
    }

    public function ndarray($x)
    {
        if($x instanceof NDArrayCL) {        // NDArrayCuda extends NDArrayCL
            $x = $x->toNDArray();
        }
        return $x;
    }
}
